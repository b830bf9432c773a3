"""GPU parity tests for row a10 beyond unit ranges: StepRange and Vector{Int} indices in ``getindex`` / ``Array(::SubDArray)``
(reference src/darray.jl:661, 706-781, 798-820), ``DArray(::SubDArray)`` (:603-609), and SubDArrays as operands of ``mapreduce`` /
broadcast (src/mapreduce.jl:36).  Data movement only: everything is bit-exact against the oracle's restatement."""
import numpy as np
import pytest

from oracle import darray_oracle as orc

pytestmark = pytest.mark.gpu


def _jl(ix, n):
    """Python index -> the oracle's 1-based index object."""
    if isinstance(ix, (int, np.integer)):
        return int(ix) % n + 1
    if isinstance(ix, slice):
        lo, hi, st = ix.indices(n)
        if st == 1:
            return (lo + 1, max(lo, hi))
        return np.arange(lo, hi, st) + 1
    v = np.asarray(ix)
    return np.where(v < 0, v + n, v) + 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64, np.bool_])
def test_strided_and_vector_views_2d(dab, rt8, dtype):
    rng = np.random.default_rng(61)
    A = (rng.standard_normal((200, 173)) * 100).astype(dtype)
    for dist in (None, (8, 1), (2, 4)):
        d = dab.distribute(A, dist=dist)
        od = orc.distribute(A, nworkers=8) if dist is None else orc.distribute(A, procs=list(range(1, 9)), dist=list(dist))
        keys = [(slice(0, 200, 3), slice(None)), (slice(5, 190, 7), slice(170, 2, -5)), ([3, 199, 0, 57, 58], slice(10, 20)),
                (slice(None, None, -1), [172, 0, 86]), (4, slice(1, 173, 2)), (slice(2, 200, 9), 100), ([7], [9]),
                (np.array([], dtype=np.int64), slice(None)), (slice(199, None, -1), slice(172, None, -1)), ([-1, -200], slice(0, 173, 50))]
        for key in keys:
            want = A[np.ix_(*[np.atleast_1d(np.arange(n)[k]) for k, n in zip(key, A.shape)])]
            want = want.reshape([s for s, k in zip(want.shape, key) if not isinstance(k, (int, np.integer))])
            got = np.asarray(d[key])
            assert got.dtype == A.dtype and got.shape == want.shape and np.array_equal(got, want), key
            ow = orc.getindex_general(od, [_jl(k, n) for k, n in zip(key, A.shape)])
            assert np.array_equal(got, ow)
        with pytest.raises(IndexError):
            d[[0, 200], :]


def test_views_of_3d_and_6d_arrays(dab, rt8):
    rng = np.random.default_rng(67)
    B = rng.standard_normal((37, 29, 5))
    d = dab.distribute(B)
    assert np.array_equal(np.asarray(d[::3, 1:20, 1]), B[::3, 1:20, 1])
    assert np.array_equal(np.asarray(d[[4, 0, 29, 16], ::-2, :]), B[np.ix_([4, 0, 29, 16], np.arange(28, -1, -2), np.arange(5))])
    # more than 4 dimensions: unit-range and strided views go through the gather kernel
    Cc = rng.integers(-50, 50, (6, 5, 4, 3, 4, 5)).astype(np.int32)
    e = dab.distribute(Cc, procs=list(range(1, 9)), dist=(2, 1, 2, 1, 2, 1))
    assert e.layout.grid == (2, 1, 2, 1, 2, 1)
    assert np.array_equal(dab.to_array(e), Cc)
    assert np.array_equal(np.asarray(e[1:5, :, 1:3, 2, ::2, 1:4]), Cc[1:5, :, 1:3, 2, ::2, 1:4])
    assert np.array_equal(np.asarray(e[:, 4, :, :, 3, :]), Cc[:, 4, :, :, 3, :])
    assert int(dab.sum(e)) == int(Cc.sum())
    r = dab.sum(e, dims=(1, 3, 5))
    assert np.array_equal(dab.to_array(r), Cc.sum(axis=(0, 2, 4), keepdims=True))


def test_darray_from_subdarray_and_view_operands(dab, rt8):
    """``s == DArray(s)`` incl. the empty view (test/darray.jl:759-771); ``sum``/``maximum``/``count`` of a view go through DArray(s)
    (src/mapreduce.jl:36); a view inside a broadcast; views of views compose."""
    rng = np.random.default_rng(71)
    A = rng.random((20, 20))
    a = dab.distribute(A)
    oa = orc.distribute(A, nworkers=8)
    s = a[0:5, 4:8]
    assert isinstance(s, dab.SubDArray)
    D = s.to_darray()
    oD = orc.darray_from_view(oa, [(1, 5), (5, 8)])
    assert D.dims == (5, 4) and list(D.layout.pids) == oD.pids and list(D.layout.indices) == oD.indices
    assert np.array_equal(dab.to_array(D), A[0:5, 4:8]) and dab.isequal(s, D) and dab.isequal(D, s)
    s0 = a[5:5, 4:8]                                                        # view(a, 6:5, 5:8): empty
    D0 = s0.to_darray()
    assert D0.dims == (0, 4) and dab.isequal(s0, D0) and dab.to_array(D0).shape == (0, 4)
    v = a[2:19:3, [17, 3, 4]]
    want = A[2:19:3][:, [17, 3, 4]]
    assert float(dab.sum(v)) == pytest.approx(want.sum(), rel=1e-14)
    assert dab.maximum(v) == want.max() and dab.minimum(v) == want.min()
    assert dab.count(v, lambda t: t > 0.5) == int((want > 0.5).sum())
    assert dab.extrema(v) == (want.min(), want.max())
    assert np.allclose(dab.to_array(dab.sum(v, dims=1)), want.sum(axis=0, keepdims=True), rtol=1e-14)
    b = dab.dzeros(want.shape)
    dab.broadcast_into(b, lambda t: 3 + t * t, v)                            # a .= 3 .+ abs2.(view(...))
    assert np.array_equal(dab.to_array(b), 3 + want * want)
    r = dab.broadcast(lambda t, u: t - u, v, want)
    assert not dab.to_array(r).any()
    vv = v[1:4, ::2]                                                         # a view of a view is a view of the parent
    assert isinstance(vv, dab.SubDArray) and vv.parent is a and np.array_equal(np.asarray(vv), want[1:4, ::2])
    assert dab.isequal(v, want) and not dab.isequal(v, want + 1)
