"""CPU tests of the Level-2 linear algebra widening (reference src/linalg.jl:1-17, 78-187, 280-311): the oracle restatement
against plain NumPy on the gathered arrays (the way the reference's own tests compare, test/darray.jl:270-282, 713-733, 933-941)
and the pure host logic of the exchange (no GPU)."""
import numpy as np
import pytest

from oracle import darray_oracle as orc


@pytest.mark.parametrize("grid,shape", [((1, 1), (5, 7)), ((2, 1), (9, 4)), ((1, 2), (6, 11)), ((2, 4), (37, 53)), ((4, 2), (20, 20)),
                                        ((3, 1), (2, 5))])
@pytest.mark.parametrize("trans", [False, True])
def test_oracle_matvec_matches_numpy(grid, shape, trans):
    rng = np.random.default_rng(5)
    nw = grid[0] * grid[1]
    for dt in (np.float64, np.float32, np.int64, np.int32):
        if np.dtype(dt).kind == "f":
            A = rng.standard_normal(shape).astype(dt)
            x = rng.standard_normal(shape[0] if trans else shape[1]).astype(dt)
        else:
            A = rng.integers(-2 ** 20, 2 ** 20, shape).astype(dt)
            x = rng.integers(-2 ** 20, 2 ** 20, shape[0] if trans else shape[1]).astype(dt)
        dA = orc.distribute(A, procs=list(range(1, nw + 1)), dist=list(grid))
        y = orc.darray_matvec(dA, x, trans)
        assert y.grid == ((grid[1],) if trans else (grid[0],))
        g0 = grid[0]
        assert y.pids == ([dA.pids[j * g0] for j in range(grid[1])] if trans else dA.pids[:g0])   # procs(A)[1,:] / procs(A)[:,1]
        got = orc.to_array(y)
        M = A.T if trans else A
        if np.dtype(dt).kind == "f":
            want = M.astype(np.float64) @ x.astype(np.float64)
            scale = np.abs(M.astype(np.float64)) @ np.abs(x.astype(np.float64)) + 1e-300
            assert np.all(np.abs(got - want) <= (1e-6 if dt == np.float32 else 1e-14) * scale)
        else:
            with np.errstate(over="ignore"):
                want = (M.astype(dt) * x[None, :]).sum(axis=1, dtype=dt)                           # wraps like Julia's Int
            assert got.dtype == dt and np.array_equal(got, want)


def test_oracle_mul_alpha_beta_and_errors():
    rng = np.random.default_rng(6)
    A = rng.integers(-5, 6, (12, 9)).astype(np.int64)
    x = rng.integers(-5, 6, 9).astype(np.int64)
    y0 = rng.integers(-5, 6, 12).astype(np.int64)
    dA = orc.distribute(A, procs=[1, 2, 3, 4, 5, 6], dist=[2, 3])
    y = orc.distribute(y0, procs=[1, 2], dist=[2])
    assert np.array_equal(orc.to_array(orc.darray_mul_vec(y, dA, x, 3, 2)), 3 * (A @ x) + 2 * y0)
    assert np.array_equal(orc.to_array(orc.darray_mul_vec(y, dA, x, 1, 1)), (A @ x) + y0)
    assert np.array_equal(orc.to_array(orc.darray_mul_vec(y, dA, x)), A @ x)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        orc.darray_mul_vec(y, dA, x[:-1])
    ybad = orc.distribute(y0, procs=[1, 2, 3], dist=[3])
    with pytest.raises(ValueError, match="ArgumentError"):
        orc.darray_mul_vec(ybad, dA, x)


@pytest.mark.parametrize("shape,nw", [((100, 200), 4), ((200, 100), 4), ((7, 3), 2), ((1, 9), 3), ((64, 64), 8)])
def test_oracle_transpose_and_diag(shape, nw):
    """reference test/darray.jl:713-733 (copy(transpose(A)) == transpose(Array(A))) and :270-282 (lmul!/rmul! with Diagonal)."""
    rng = np.random.default_rng(7)
    A = rng.standard_normal(shape)
    dA = orc.distribute(A, nworkers=nw)
    T = orc.darray_transpose(dA)
    assert T.dims == (shape[1], shape[0]) and T.pids == dA.pids[:len(T.pids)]
    assert T.grid == tuple(orc.defaultdist_grid(T.dims, len(dA.pids)))                  # default dist over procs(D)
    assert np.array_equal(orc.to_array(T), A.T)
    b = rng.standard_normal(shape[0])
    assert np.array_equal(orc.to_array(orc.darray_scale_diag(dA, b, "l")), b[:, None] * A)
    c = rng.standard_normal(shape[1])
    assert np.array_equal(orc.to_array(orc.darray_scale_diag(dA, c, "r")), A * c[None, :])


@pytest.mark.parametrize("grid", [(2, 4), (4, 2), (8, 1), (1, 8), (2, 2), (3, 2)])
@pytest.mark.parametrize("trans", [False, True])
@pytest.mark.parametrize("wpr", [1, 2])
def test_matvec_exchange_plan_is_matched(grid, trans, wpr):
    """Every rank derives its sends/recvs from the layouts alone; for each ordered rank pair the sender's list and the receiver's
    list must be the same (i, j, length) sequence -- the NCCL matching rule for grouped point-to-point calls."""
    from darray_b200._linalg import matvec_exchange_plan
    from darray_b200.layout import make_layout
    nw = grid[0] * grid[1]
    pids = list(range(2, nw + 2))
    dims = (37, 53)
    L = make_layout(dims, pids, list(grid))
    rd = 1 if trans else 0
    g0 = grid[0]
    ypids = [pids[j * g0] for j in range(grid[1])] if trans else pids[:g0]
    yl = make_layout((dims[rd],), ypids, [grid[rd]])
    rank_of = lambda pid: (pid - 2) // wpr                                                  # noqa: E731
    nranks = (nw + wpr - 1) // wpr
    plans = [matvec_exchange_plan(L, yl, trans, rank_of, r) for r in range(nranks)]
    gi, gj = (grid[1], grid[0]) if trans else grid
    covered = set()
    for a in range(nranks):
        for b in range(nranks):
            s = [(i, j, n) for i, j, n, peer in plans[a]["sends"] if peer == b]
            r = [(i, j, n) for i, j, n, peer in plans[b]["recvs"] if peer == a]
            assert s == r
            covered.update((i, j) for i, j, _ in s)
        covered.update((i, j) for i, j, _ in plans[a]["local"])
        assert all(rank_of(yl.pids[i]) == a for i in plans[a]["owned"])
    assert covered == {(i, j) for i in range(gi) for j in range(gj)}
    assert sorted(i for p in plans for i in p["owned"]) == list(range(gi))
