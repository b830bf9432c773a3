"""Pins the CPU oracle (oracle/darray_oracle.py) to every known-answer vector the reference's own tests / docs hold for the
hot path.  The reference compares DArray ops against the same Base op on the gathered Array; here the
"gathered Array" side is NumPy (exact for ints / elementwise), the "DArray" side is the oracle's restatement of the
reference's distributed algorithm (layout -> per-chunk op -> combine).  Citations: reference test/darray.jl unless noted.
"""
import numpy as np
import pytest

from oracle import darray_oracle as orc


# ------------------------------------------------------------------------------------------------ layout KATs
def test_defaultdist_50_4():
    """test/darray.jl:66  DistributedArrays.defaultdist(50,4) == [1,14,27,39,51]"""
    assert orc.defaultdist_cuts(50, 4) == [1, 14, 27, 39, 51]


def test_uneven_distribution_issue_166():
    """test/darray.jl:62-64: drand((3,), [MYID, OTHERIDS]) -> first worker gets 2, second 1."""
    d = orc.make_layout((3,), [7, 9])
    assert [orc.rlen(i[0]) for i in d.indices] == [2, 1] and d.pids == [7, 9]


def test_grids_of_baseline_configs():
    """C1 -> (1,2); C3 -> (8,); C4 -> (2,4) (ties go to the highest dim, src/darray.jl:266-268)."""
    assert orc.defaultdist_grid((1024, 1024), 2) == [1, 2]
    assert orc.defaultdist_grid((8 << 30,), 8) == [8]
    assert orc.defaultdist_grid((65536, 65536), 8) == [2, 4]
    assert orc.defaultdist_grid((20, 20, 20), 8) == [2, 2, 2]
    assert orc.defaultdist_grid((100, 100), 6) == [2, 3]
    assert orc.defaultdist_grid((2, 100), 8) == [1, 8]


def test_chunk_origin_200x200():
    """test/darray.jl:193-194: for 200x200 over 2 procs, D[1,101] is the second worker's localpart[1,1]."""
    A = np.arange(200 * 200, dtype=np.float64).reshape((200, 200), order="F")
    d = orc.distribute(A, procs=[1, 2])
    assert d.grid == (1, 2) and d.indices == [((1, 200), (1, 100)), ((1, 200), (101, 200))]
    assert d.chunks[0][0, 0] == A[0, 0] and d.chunks[1][0, 0] == A[0, 100]
    assert orc.locate(d.cuts, (1, 101)) == (1, 2) and orc.locate(d.cuts, (200, 100)) == (1, 1)
    with pytest.raises(ValueError):
        orc.locate(d.cuts, (1, 201))


def test_roundtrip_and_irregular_chunks():
    """test/darray.jl:225-231 (copy! with 3+7-row chunks) and :306-310: DArray(refs) layout from chunk sizes."""
    r1, r2 = np.ones((3, 10)), 2 * np.ones((7, 10))
    d = orc.from_chunks([r1, r2], (2, 1), [2, 3])
    assert d.dims == (10, 10) and d.indices == [((1, 3), (1, 10)), ((4, 10), (1, 10))] and d.cuts == [[1, 4, 11], [1, 11]]
    assert np.array_equal(orc.to_array(d), np.vstack([r1, r2]))
    for shape in [(50,), (1024, 1024), (73, 73), (20, 20, 20), (7, 1), (2, 3, 5, 4)]:
        A = np.arange(int(np.prod(shape)), dtype=np.int64).reshape(shape, order="F")
        assert np.array_equal(orc.to_array(orc.distribute(A, nworkers=8)), A)


# ------------------------------------------------------------------------------------------------ reduce / map / map!
def test_reduce_fill_myid():
    """test/darray.jl:238-257: D = fill(myid()) on (10,10) over [MYID, OTHERIDS]."""
    MY, OT = 1, 5
    d = orc.make_layout((10, 10), [MY, OT])
    d.chunks = [np.full([orc.rlen(r) for r in idx], pid, dtype=np.int64, order="F") for pid, idx in zip(d.pids, d.indices)]
    assert orc.darray_mapreduce(None, "+", d)[0] == 50 * MY + 50 * OT
    d2 = orc.darray_broadcast(lambda x: np.ones_like(x), d, d)
    assert orc.darray_mapreduce(None, "+", d2)[0] == 100
    orc.darray_map_inplace(lambda x: np.ones_like(x), d, d)
    assert orc.darray_mapreduce(None, "+", d)[0] == 100


def test_mapreduce_int_exact():
    """test/darray.jl:286-294: mapreduce(f, opt, DA) - mapreduce(f, opt, A) == 0 for integer-valued f, opt in (+, *).
    The reference maps to Int128 so that products cannot overflow; exact Python ints (object dtype) play that role here."""
    rng = np.random.default_rng(0)
    fs = [lambda x: 2 * x, lambda x: x * x, lambda x: x * x + 2 * x - 1]
    for _ in range(25):
        for f in fs:
            for op, red in (("+", np.sum), ("*", np.prod)):
                A = rng.integers(1, 6, rng.integers(2, 31)).astype(object)
                d = orc.distribute(A, nworkers=4)
                assert orc.darray_mapreduce(f, op, d)[0] - red(f(A)) == 0
                # the Int128 restatement (machine arithmetic: wraps at 128 bits) agrees with the exact integers wherever they fit, and
                # wraps like Julia where they do not (x^2 + 2x - 1 = 34 at x = 5: 34^30 > 2^127)
                d64 = orc.distribute(A.astype(np.int64), nworkers=4)
                assert orc.darray_mapreduce_int128(f, op, d64) == orc.wrap_int128(int(red(f(A))))
    full = orc.distribute(np.full(30, 5, dtype=np.int64), nworkers=4)
    got = orc.darray_mapreduce_int128(fs[2], "*", full)
    assert got == orc.wrap_int128(34 ** 30) and got != 34 ** 30 and -2 ** 127 <= got < 2 ** 127


def test_max_min_sum_int():
    """test/darray.jl:439-452."""
    rng = np.random.default_rng(1)
    a = (np.round(rng.random((100, 1000)) * 100) - 50).astype(np.int64)
    d = orc.distribute(a, nworkers=5)
    mr = lambda f, op: orc.darray_mapreduce(f, op, d)[0]
    assert mr(None, "+") == a.sum() and mr(None, "max") == a.max() and mr(None, "min") == a.min()
    assert mr(np.abs, "max") == np.abs(a).max() and mr(np.abs, "min") == np.abs(a).min()
    assert mr(np.abs, "+") == np.abs(a).sum() and mr(lambda x: x * x, "+") == (a * a).sum()
    assert orc.darray_extrema(d) == (a.min(), a.max())


def test_all_any_count_prod():
    """test/darray.jl:456-518."""
    a = np.ones(100, dtype=bool)
    ident = lambda x: x
    d = orc.distribute(a, nworkers=4)
    assert orc.darray_all(ident, d) and orc.darray_any(ident, d)
    a[49] = False
    d = orc.distribute(a, nworkers=4)
    assert not orc.darray_all(ident, d) and orc.darray_any(ident, d)
    d = orc.distribute(np.zeros(100, dtype=bool), nworkers=4)
    assert not orc.darray_all(ident, d) and not orc.darray_any(ident, d)
    o = np.ones((10, 10))
    o[9, 0] = 2.0
    d = orc.distribute(o, nworkers=4)
    assert orc.darray_count(lambda x: x == 2.0, d) == 1 and orc.darray_count(lambda x: x == 1.0, d) == 99
    assert orc.darray_count(lambda x: x == 0.0, d) == 0
    assert orc.darray_any(lambda x: x == 2.0, d) and not orc.darray_any(lambda x: x == 3.0, d)
    assert orc.darray_mapreduce(None, "*", orc.distribute(np.full(10, 2, dtype=np.int64), nworkers=3))[0] == 2 ** 10


# ------------------------------------------------------------------------------------------------ mapreducedim
def test_mapreducedim_73x73_int():
    """test/darray.jl:298-304: ones(73,73) over 2 procs (37/36 split), t->t*t, +, dims = 1, 2, (1,2): exact."""
    d = orc.make_layout((73, 73), [1, 2])
    d.chunks = [np.ones([orc.rlen(r) for r in idx], dtype=np.int64, order="F") for idx in d.indices]
    A = np.ones((73, 73), dtype=np.int64)
    for region, ax in (([1], 0), ([2], 1), ([1, 2], (0, 1))):
        R = orc.darray_mapreducedim(lambda t: t * t, "+", d, region)
        assert np.array_equal(orc.to_array(R), (A * A).sum(axis=ax, keepdims=True))
    R = orc.darray_mapreducedim(None, "+", d, [2])
    assert R.grid == (1, 1) and R.pids == [1]  # "Store reduction on lowest pids" (src/mapreduce.jl:43-44)
    R = orc.darray_mapreducedim(None, "+", d, [1])
    assert R.grid == (1, 2) and R.pids == [1, 2] and R.indices == [((1, 1), (1, 37)), ((1, 1), (38, 73))]


def test_irregular_sum_dims2():
    """test/darray.jl:306-310."""
    rng = np.random.default_rng(2)
    r1, r2 = rng.standard_normal((3, 10)), rng.standard_normal((7, 10))
    D = orc.from_chunks([r1, r2], (2, 1), [1, 2])
    A = np.vstack([r1, r2])
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "+", D, [2])), A.sum(axis=1, keepdims=True), rtol=1e-14)


@pytest.mark.parametrize("dms", [(1,), (2,), (3,), (1, 2), (1, 3), (2, 3), (1, 2, 3)])
def test_mapreducedim_all_subsets_float64(dms):
    """test/darray.jl:319-332: 20^3 randn, with and without init, `≈` (rtol sqrt(eps))."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((20, 20, 20))
    d = orc.distribute(A, nworkers=8)
    ax = tuple(x - 1 for x in dms)
    rtol = float(np.sqrt(np.finfo(np.float64).eps))
    sq = lambda t: t * t
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(sq, "+", d, dms)), (A * A).sum(axis=ax, keepdims=True), rtol=rtol)
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(sq, "+", d, dms, init=1.0)), (A * A).sum(axis=ax, keepdims=True) + 1.0, rtol=rtol)
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "*", d, dms)), A.prod(axis=ax, keepdims=True), rtol=rtol)
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "*", d, dms, init=2.0)), 2.0 * A.prod(axis=ax, keepdims=True), rtol=rtol)


def test_sum_dims_errors_and_noop_dim():
    """test/darray.jl:357-401: dims=-1 / 0 -> ArgumentError; dims=3 on a matrix is a no-op dimension."""
    A = np.random.default_rng(4).standard_normal((100, 100))
    d = orc.distribute(A, nworkers=8)
    for bad in (-1, 0):
        with pytest.raises(ValueError):
            orc.darray_mapreducedim(None, "+", d, [bad])
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "+", d, [1])), A.sum(axis=0, keepdims=True), rtol=1e-12)
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "+", d, [2])), A.sum(axis=1, keepdims=True), rtol=1e-12)
    assert np.allclose(orc.to_array(orc.darray_mapreducedim(None, "+", d, [3])), A, rtol=1e-12)
    assert np.isclose(orc.darray_mapreduce(None, "+", d)[0], A.sum(), rtol=1e-12)


# ------------------------------------------------------------------------------------------------ halo reads / makelocal
def test_subdarray_to_array():
    """test/darray.jl:185-218."""
    A = np.random.default_rng(5).random((200, 200))
    D = orc.distribute(A, procs=[1, 2])
    get = lambda *J: orc.getindex_array(D, J)
    assert np.array_equal(get((1, 150), (1, 150)), A[0:150, 0:150])
    assert np.array_equal(get((4, 4), (23, 176))[0], A[3, 22:176])
    assert np.array_equal(get((23, 176), (197, 197))[:, 0], A[22:176, 196])
    assert np.array_equal(get((3, 4), (99, 100)), A[2:4, 98:100])
    assert np.array_equal(get((1, 1), (1, 4))[0], A[0, 0:4])
    plan = orc.slab_plan(D, ((3, 4), (99, 102)))
    assert [(c, src, dst, whole) for c, src, dst, whole in plan] == [(0, ((3, 4), (99, 100)), ((1, 2), (1, 2)), False),
                                                                       (1, ((3, 4), (1, 2)), ((1, 2), (3, 4)), False)]
    assert orc.slab_plan(D, ((1, 200), (101, 200)))[0][3] is True  # exactly one chunk -> chunk(d, pid) (src/darray.jl:588-592)


def test_makelocal():
    """test/darray.jl:740-757."""
    n = 45
    A = np.random.default_rng(6).standard_normal((n, n))
    dA = orc.distribute(A, nworkers=9)
    for i in range(n):
        assert np.array_equal(orc.getindex_array(dA, ((1, n), (i + 1, i + 1)))[:, 0], A[:, i])
        assert np.array_equal(orc.getindex_array(dA, ((i + 1, i + 1), (1, n)))[0, :], A[i, :])
    assert np.array_equal(orc.getindex_array(dA, ((1, 5), (1, 5))), A[:5, :5])
    pid = dA.pids[3]
    lid = orc.localindices(dA, pid)
    assert orc.makelocal_view_ranges(dA, pid, lid) == tuple((1, orc.rlen(r)) for r in lid)   # zero-copy branch
    assert orc.makelocal_view_ranges(dA, pid, ((1, n), (1, 1))) is None                       # halo-fetch branch


# ------------------------------------------------------------------------------------------------ broadcast
def test_broadcast_ops():
    """test/darray.jl:880-905 and :845-856 (scalar ops)."""
    nw = 4
    nrows, ncols = 20 * nw, 10 * nw
    rng = np.random.default_rng(7)
    A = rng.random((nrows, ncols))
    a = orc.distribute(A, procs=list(range(1, nw + 1)), dist=(1, nw))
    M = A.mean(axis=0, keepdims=True)
    m = orc.distribute(M, nworkers=nw)
    c = orc.darray_broadcast(lambda x, y: x - y, orc.make_layout(A.shape, list(range(1, nw + 1))), a, m)
    assert np.array_equal(orc.to_array(c), A - M)
    g = orc.darray_broadcast(lambda x, y, z: x - y * np.sin(z), orc.make_layout(A.shape, list(range(1, nw + 1))), a, m, c)
    assert np.array_equal(orc.to_array(g), A - M * np.sin(A - M))
    r = orc.darray_broadcast(lambda o: o, a, np.ones((nrows, ncols)))
    assert (orc.to_array(r) == 1).all()
    Z = np.zeros((nrows, ncols + 5))[:, 5:]
    r = orc.darray_broadcast(lambda z: 3 + z * z, a, Z)
    assert (orc.to_array(r) == 3).all()
    B = rng.random((20, 20))
    C = rng.random((20, 20))
    b, cc = orc.distribute(B, nworkers=nw), orc.distribute(C, nworkers=nw)
    x = rng.random()
    lay = orc.make_layout((20, 20), [1, 2, 3, 4])
    for f in (np.add, np.subtract, np.multiply, np.divide, np.fmod):
        assert np.array_equal(orc.to_array(orc.darray_broadcast(lambda u: f(u, x), lay, b)), f(B, x))
        assert np.array_equal(orc.to_array(orc.darray_broadcast(lambda u: f(x, u), lay, b)), f(x, B))
        assert np.array_equal(orc.to_array(orc.darray_broadcast(f, lay, b, cc)), f(B, C))


def test_affine_is_two_roundings():
    x = orc.rand_u01(7, 0, 1 << 14) + np.float32(1)
    a, b = np.float32(1.0000001), np.float32(-1.0000001)
    got = orc.affine_unfused(a, x, b)
    fused = (x.astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)
    assert got.dtype == np.float32 and (got != fused).any()
    assert np.array_equal(got, (a * x).astype(np.float32) + b)


# ------------------------------------------------------------------------------------------------ float reduction order (docs goldens)
def test_docs_golden_local_sum():
    """docs/src/index.md:222-225: sum(fill(1.1,(100,100))) == 11000.000000000013 -- pins the pairwise-1024 + 16-accumulator
    (AVX2 Float64: 4 lanes x 4 interleave) model of Base.mapreduce_impl."""
    A = np.full((100, 100), 1.1)
    assert repr(float(orc.julia_mapreduce(None, "+", A))) == "11000.000000000013"


def test_docs_golden_distributed_sum_is_layout_dependent():
    """docs/src/index.md:227-236: the distributed sum differs from the local one and depends on the layout.  The documented
    value 11000.000000000127 was recorded on an older Julia/layout: it is reproduced by a 1-D split over 8 workers with a
    scalar (non-SIMD) base block; today's default layout (2x4 grid, 16-accumulator base block) gives ...013 again."""
    A = np.full((100, 100), 1.1)
    d = orc.distribute(A, procs=list(range(1, 9)), dist=(8, 1))
    assert repr(float(orc.darray_mapreduce(None, "+", d, simd=(1, 1))[0])) == "11000.000000000127"
    d = orc.distribute(A, nworkers=8)
    assert d.grid == (2, 4)
    assert repr(float(orc.darray_mapreduce(None, "+", d)[0])) == "11000.000000000013"


def test_julia_max_min_semantics():
    """NaN-propagating, maximum([0.0,-0.0]) === 0.0, minimum([0.0,-0.0]) === -0.0, empty throws."""
    f = np.float32
    assert np.isnan(orc.julia_mapreduce(None, "max", np.array([1, np.nan, 3], dtype=f)))
    assert np.isnan(orc.julia_mapreduce(None, "min", np.array([1, 2, np.nan] * 10, dtype=f)))
    z = orc.julia_mapreduce(None, "max", np.array([0.0, -0.0], dtype=f))
    assert z == 0 and not np.signbit(z)
    z = orc.julia_mapreduce(None, "min", np.array([0.0, -0.0], dtype=f))
    assert z == 0 and np.signbit(z)
    with pytest.raises(ValueError):
        orc.julia_mapreduce(None, "max", np.zeros(0, dtype=f))
    assert orc.julia_mapreduce(None, "+", np.zeros(0, dtype=f)) == 0
    s = orc.julia_mapreduce(None, "+", np.array([1, 2, 3], dtype=np.int32))
    assert s.dtype == np.int64  # add_sum widens Int32
    assert orc.julia_mapreduce(None, "+", np.ones(5, dtype=f)).dtype == np.float32


def test_c1_fixture_map_and_sum():
    """BASELINE config 1: distribute(rand(Float32,1024,1024)) on 2 workers; map!(x->2x+1) and sum."""
    A = orc.rand_u01(1234, 0, 1 << 20).reshape((1024, 1024), order="F")
    d = orc.distribute(A, nworkers=2)
    assert d.grid == (1, 2) and [c.shape for c in d.chunks] == [(1024, 512), (1024, 512)]
    orc.darray_map_inplace(lambda x: np.float32(2) * x + np.float32(1), d, d)
    want = np.float32(2) * A + np.float32(1)
    assert np.array_equal(orc.to_array(d), want)
    s, parts = orc.darray_mapreduce(None, "+", d)
    assert s.dtype == np.float32 and s == np.float32(parts[0] + parts[1])
    exact = 2 * orc.rand_u01_ksum(1234, 0, 1 << 20) * 2.0 ** -24 + (1 << 20)
    assert abs(float(s) - exact) <= 1e-6 * exact


def test_oracle_general_views_match_numpy_and_reference_slices():
    """getindex_general (StepRange / Vector{Int} / Int indices, src/darray.jl:661, 798-820) against NumPy fancy indexing on the gathered
    array -- the comparison every reference view test makes (test/darray.jl:185-218, 740-757) -- including the reference's own slices."""
    rng = np.random.default_rng(3)
    A = rng.standard_normal((200, 200))
    d = orc.distribute(A, procs=[1, 2], dist=[1, 2])
    assert np.array_equal(orc.getindex_general(d, [(1, 150), (1, 150)]), A[:150, :150])          # D[1:150, 1:150]
    assert np.array_equal(orc.getindex_general(d, [4, (23, 176)]), A[3, 22:176])                  # D[4, 23:176]
    assert np.array_equal(orc.getindex_general(d, [(23, 176), 197]), A[22:176, 196])              # D[23:176, 197]
    B = rng.standard_normal((37, 29, 5))
    e = orc.distribute(B, nworkers=8)
    for I, np_ix in [([np.arange(1, 38, 3), (2, 20), 2], np.ix_(np.arange(0, 37, 3), np.arange(1, 20), [1])),
                     ([np.array([5, 1, 30, 17]), np.arange(29, 0, -2), (1, 5)], np.ix_([4, 0, 29, 16], np.arange(28, -1, -2), np.arange(5))),
                     ([(1, 37), np.array([], dtype=np.int64), (1, 5)], np.ix_(np.arange(37), [], np.arange(5)))]:
        got = orc.getindex_general(e, I)
        want = B[np_ix]
        want = want.reshape([n for n, ix in zip(want.shape, I) if not isinstance(ix, (int, np.integer))])
        assert got.shape == want.shape and np.array_equal(got, want)
    v = orc.darray_from_view(d, [(1, 5), (5, 8)])                                                 # test/darray.jl:759-771
    assert v.dims == (5, 4) and np.array_equal(orc.to_array(v), A[:5, 4:8])
