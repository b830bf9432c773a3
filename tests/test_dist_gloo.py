"""world_size-2 tests on CPU over gloo: the host-side logic of the N > 1 path.

The data path of the product is GPU-only (NCCL / peer loads through libdab200.so); what CAN and must be checked without a GPU is
that every rank derives the same global plan and that the plans match pairwise: the sends one rank issues are exactly the
receives its peer posts, in the same order (the grouped ncclSend/ncclRecv matching rule), and that executing the plan reproduces
the oracle's ``mapreducedim`` / ``reduce(op, results)``.  The executor below lives in the TEST (NumPy + gloo send/recv); the
per-chunk partials come from the oracle.  The product code exercised: layout.make_layout, _mapreduce.plan_reducedim,
_mapreduce.exchange_plan, _mapreduce._fold (dab_combine_ordered, host-only C entry point).
"""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, wpr, q):
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist

        import darray_b200 as dab
        from darray_b200._mapreduce import _fold, exchange_plan, plan_reducedim
        from darray_b200 import _lib
        from oracle import darray_oracle as orc

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        P = world * wpr
        rank_of = lambda pid: (pid - 1) // wpr
        rng = np.random.default_rng(42)  # same data on every rank (the test's stand-in for @everywhere)
        for shape, region in [((12, 10), (1,)), ((12, 10), (2,)), ((12, 10), (1, 2)), ((6, 5, 8), (1, 3)), ((6, 5, 8), (2,)), ((9,), (1,))]:
            A = rng.integers(-9, 9, shape).astype(np.int64)
            procs = list(range(1, min(P, max(shape)) + 1))
            L = dab.make_layout(shape, procs)
            od = orc.distribute(A, procs=procs)
            assert L.indices == od.indices and L.grid == tuple(od.grid)
            Rlayout, fibres = plan_reducedim(L, region)
            want = orc.darray_mapreducedim(None, "+", od, region)
            assert Rlayout.pids == want.pids and Rlayout.indices == want.indices and Rlayout.grid == tuple(want.grid) and Rlayout.cuts == want.cuts
            # every rank holds only its own workers' partials (phase 1 result, from the oracle)
            mine = {pid: orc.julia_mapreducedim(None, "+", od.chunks[i], region) for i, pid in enumerate(L.pids) if rank_of(pid) == rank}
            xp = exchange_plan(L, Rlayout, fibres, rank_of, rank)
            # (1) plans agree pairwise: my sends == what my peers expect to receive from me, in order
            allxp = [None] * world
            dist.all_gather_object(allxp, xp)
            for peer in range(world):
                if peer == rank:
                    continue
                my_sends = [(mp, rl) for mp, dst, rl in xp["sends"] if dst == peer]
                their_recvs = [(mp, rl) for rl, slot, mp, src in allxp[peer]["recvs"] if src == rank]
                assert my_sends == their_recvs, (shape, region, rank, peer)
            # (2) execute the plan over gloo, accumulate in fibre order, compare with the oracle's R
            stacks = {rl: [None] * len(fibres[rl]) for rl in xp["owned"]}
            for rl, slot, mp in xp["local"]:
                stacks[rl][slot] = mine[mp]
            reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(mine[mp].ravel(order="F"))), dst) for mp, dst, rl in xp["sends"]]
            for rl, slot, mp, src in xp["recvs"]:
                shp = tuple(orc.rlen(r) for r in Rlayout.indices[rl])
                buf = torch.empty(int(np.prod(shp)), dtype=torch.int64)
                dist.recv(buf, src)
                stacks[rl][slot] = buf.numpy().reshape(shp, order="F")
            for r in reqs:
                r.wait()
            for rl, parts in stacks.items():
                R = np.zeros_like(parts[0])
                for p in parts:
                    R = R + p
                i = want.pids.index(Rlayout.pids[rl])
                assert np.array_equal(R, want.chunks[i]), (shape, region, rl)
            dist.barrier()
        # (3) reduce(op, results): every rank contributes its workers' chunk results; all ranks fold identically, in procs order
        x = orc.rand_u01(7, 0, 100003)
        procs = list(range(1, P + 1))
        L = dab.make_layout(x.shape, procs)
        od = orc.distribute(x, procs=procs)
        host = np.zeros(16 * P, dtype=np.uint8)
        local = {pid: orc.julia_mapreduce(None, "+", od.chunks[i]) for i, pid in enumerate(L.pids) if rank_of(pid) == rank}
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        for g in gathered:
            for pid, v in g.items():
                host[16 * (pid - 1):16 * (pid - 1) + 4] = np.asarray([v], dtype=np.float32).view(np.uint8)
        res, vals = _fold(host, L.pids, np.dtype(np.float32), _lib.SUM)
        ref, parts = orc.darray_mapreduce(None, "+", od)
        assert res == ref and list(vals) == [np.float32(p) for p in parts]
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("wpr", [1, 2])
def test_world2_exchange_plans_and_ordered_fold(wpr):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, wpr, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
