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


def _worker_level2_sort(rank, world, port, wpr, q):
    """mul!(y, A, x) and sort(d) across 2 ranks: the product's plans (``_linalg.matvec_exchange_plan``, ``_sort.sort_exchange_plan``,
    ``_sort.boundaries_from_samples``) executed with NumPy + gloo in the TEST; per-chunk work (tile products, chunk sorts, split
    points) comes from the oracle.  The result must equal the oracle's distributed restatement chunk by chunk."""
    try:
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist

        import darray_b200 as dab
        from darray_b200._linalg import matvec_exchange_plan
        from darray_b200._sort import boundaries_from_samples, sort_exchange_plan
        from oracle import darray_oracle as orc

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        P = world * wpr
        rank_of = lambda pid: (pid - 1) // wpr                                                  # noqa: E731
        rng = np.random.default_rng(7)
        # ---- mul!(y, A, x, 3, 2) and its transposed form
        for grid in [(P, 1), (1, P)] + ([(2, 2)] if P == 4 else []):
            for trans in (False, True):
                A = rng.integers(-9, 9, (23, 17)).astype(np.int64)
                x = rng.integers(-9, 9, 23 if trans else 17).astype(np.int64)
                pids = list(range(1, P + 1))
                L = dab.make_layout(A.shape, pids, list(grid))
                oA = orc.distribute(A, procs=pids, dist=list(grid))
                rd, cd = (1, 0) if trans else (0, 1)
                g0 = grid[0]
                ypids = [pids[j * g0] for j in range(grid[1])] if trans else pids[:g0]
                y0 = rng.integers(-9, 9, A.shape[rd]).astype(np.int64)
                yl = dab.make_layout((A.shape[rd],), ypids, [grid[rd]])
                oy = orc.distribute(y0, procs=ypids, dist=[grid[rd]])
                want = orc.darray_mul_vec(oy, oA, x, 3, 2, trans)
                gi, gj = (grid[1], grid[0]) if trans else grid
                plan = matvec_exchange_plan(L, yl, trans, rank_of, rank)
                tile = {}
                for i in range(gi):
                    for j in range(gj):
                        lin = (j + i * g0) if trans else (i + j * g0)
                        if rank_of(L.pids[lin]) == rank:
                            lo, hi = L.cuts[cd][j], L.cuts[cd][j + 1] - 1
                            tile[(i, j)] = orc._tile_matvec(oA.chunks[lin], x[lo - 1:hi], trans)
                stacks = {i: [None] * gj for i in plan["owned"]}
                for i, j, plen in plan["local"]:
                    stacks[i][j] = tile[(i, j)]
                reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(tile[(i, j)])), peer) for i, j, plen, peer in plan["sends"]]
                for i, j, plen, peer in plan["recvs"]:
                    buf = torch.empty(plen, dtype=torch.int64)
                    dist.recv(buf, peer)
                    stacks[i][j] = buf.numpy()
                for r in reqs:
                    r.wait()
                for i, parts in stacks.items():
                    yi = 2 * oy.chunks[i]
                    for p in parts:
                        yi = yi + 3 * p
                    assert np.array_equal(yi, want.chunks[i]), (grid, trans, i)
                dist.barrier()
        # ---- mul!(C, A, B, 3, 2) and its transposed form: tile results R[i,j,k] travel to the owners of C's chunks (i, k)
        from darray_b200._linalg import matmat_exchange_plan
        for grid in [(P, 1), (1, P)] + ([(2, 2)] if P == 4 else []):
            for trans in (False, True):
                A = rng.integers(-9, 9, (23, 17)).astype(np.int64)
                rd, cd = (1, 0) if trans else (0, 1)
                Bm = rng.integers(-9, 9, (A.shape[cd], 11)).astype(np.int64)
                pids = list(range(1, P + 1))
                L = dab.make_layout(A.shape, pids, list(grid))
                oA = orc.distribute(A, procs=pids, dist=list(grid))
                oB = orc.distribute(Bm, procs=pids, dist=[1, P])
                oC = orc.darray_matmat(oA, oB, trans)                                          # layout of A*B as the reference builds it
                C0 = rng.integers(-9, 9, oC.dims).astype(np.int64)
                oC0 = orc.distribute(C0, procs=oC.pids, dist=list(oC.grid))
                want = orc.darray_mul_mat(oC0, oA, Bm, 3, 2, trans)
                CL = dab.make_layout(oC.dims, oC.pids, list(oC.grid))
                assert CL.indices == oC.indices
                g0 = grid[0]
                gi, gj = (grid[1], grid[0]) if trans else grid
                gk = oC.grid[1]
                plan = matmat_exchange_plan(L, CL, trans, rank_of, rank)
                tile = {}
                for i in range(gi):
                    for j in range(gj):
                        lin = (j + i * g0) if trans else (i + j * g0)
                        if rank_of(L.pids[lin]) == rank:
                            lo, hi = L.cuts[cd][j], L.cuts[cd][j + 1] - 1
                            for k in range(gk):
                                clo, chi = CL.cuts[1][k], CL.cuts[1][k + 1] - 1
                                tile[(i, j, k)] = orc._tile_matmat(oA.chunks[lin], Bm[lo - 1:hi, clo - 1:chi], trans)
                stacks = {ik: [None] * gj for ik in plan["owned"]}
                for i, j, k, rows, cols in plan["local"]:
                    stacks[(i, k)][j] = tile[(i, j, k)]
                reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(tile[(i, j, k)].reshape(-1, order="F"))), peer)
                        for i, j, k, rows, cols, peer in plan["sends"]]
                for i, j, k, rows, cols, peer in plan["recvs"]:
                    buf = torch.empty(rows * cols, dtype=torch.int64)
                    dist.recv(buf, peer)
                    stacks[(i, k)][j] = buf.numpy().reshape((rows, cols), order="F")
                for r in reqs:
                    r.wait()
                for (i, k), parts in stacks.items():
                    lin_c = i + k * CL.grid[0]
                    ci = 2 * oC0.chunks[lin_c]
                    for p in parts:
                        ci = ci + 3 * p
                    assert np.array_equal(ci, want.chunks[lin_c]), (grid, trans, i, k)
                dist.barrier()
        # ---- sort(d; sample=true)
        for n in (P, 1000, 30011):
            a = rng.integers(-10 ** 6, 10 ** 6, n).astype(np.int64)
            pids = list(range(1, P + 1))
            od = orc.distribute(a, procs=pids)
            want, wantb = orc.darray_sort(od, True)
            mine = {pid: orc.jl_sort(od.chunks[k]) for k, pid in enumerate(od.pids) if rank_of(pid) == rank}
            samples = {pid: s[list(orc.sort_sample_indices(len(s)))] for pid, s in mine.items()}
            gathered = [None] * world
            dist.all_gather_object(gathered, samples)
            allsamples = {}
            for g in gathered:
                allsamples.update(g)
            b = boundaries_from_samples(np.concatenate([allsamples[p] for p in od.pids]), len(od.pids), np.dtype(np.int64))
            assert np.array_equal(b, wantb)
            ends = {pid: orc.sort_split_points(s, b) for pid, s in mine.items()}
            sizes_mine = {pid: [e[0]] + [e[k] - e[k - 1] for k in range(1, len(e))] for pid, e in ends.items()}
            dist.all_gather_object(gathered, sizes_mine)
            sizes = {}
            for g in gathered:
                sizes.update(g)
            plan = sort_exchange_plan(od.pids, sizes, rank_of, rank)
            allplans = [None] * world
            dist.all_gather_object(allplans, plan)
            for peer in range(world):
                if peer != rank:
                    assert [(j, p, m) for j, p, m, dst in plan["sends"] if dst == peer] == \
                           [(j, p, m) for j, p, off, m, src in allplans[peer]["recvs"] if src == rank]
            totals = [sum(sizes[p][j] for p in od.pids) for j in range(len(od.pids))]
            recv = {j: np.empty(totals[j], dtype=np.int64) for j, pid in enumerate(od.pids) if rank_of(pid) == rank}
            for j, p, off, m in plan["local"]:
                recv[j][off:off + m] = mine[p][ends[p][j] - m:ends[p][j]]
            reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(mine[p][ends[p][j] - m:ends[p][j]])), peer) for j, p, m, peer in plan["sends"]]
            for j, p, off, m, peer in plan["recvs"]:
                buf = torch.empty(m, dtype=torch.int64)
                dist.recv(buf, peer)
                recv[j][off:off + m] = buf.numpy()
            for r in reqs:
                r.wait()
            for j, buf in recv.items():
                pid = od.pids[j]
                if totals[j]:
                    assert np.array_equal(np.sort(buf), want.chunks[want.pids.index(pid)]), (n, j)
                else:
                    assert pid not in want.pids
            dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("wpr", [1, 2])
def test_world2_matvec_and_sort_exchange(wpr):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_level2_sort, args=(r, 2, port, wpr, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
