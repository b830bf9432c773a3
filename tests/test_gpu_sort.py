"""GPU parity tests for K11 (dab_sort / dab_sorted_split) and the DVector samplesort (reference src/sort.jl; reference test
test/darray.jl:1015-1025).  Everything here is integer / bit-pattern work: results must equal the oracle's exactly, including the
boundaries, the chunk sizes and which workers hold the result."""
import ctypes as C

import numpy as np
import pytest

from oracle import darray_oracle as orc

pytestmark = pytest.mark.gpu


def _data(T, n, rng, kind="full"):
    if np.dtype(T).kind == "i":
        if kind == "small":
            return rng.integers(-1000, 1000, n).astype(T)                       # most digit passes are constant and skipped
        return rng.integers(np.iinfo(T).min, np.iinfo(T).max, n, dtype=T)
    if kind == "small":
        return rng.random(n).astype(T)
    return (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)).astype(T)


def _sort_chunk(dab, rt, a, inplace=False):
    from darray_b200 import _lib
    src = dab.B200Array.from_numpy(rt, a) if a.size else dab.B200Array.empty(rt, (0,), a.dtype)
    out = src if inplace else dab.B200Array.empty(rt, (a.size,), a.dtype)
    tmp = dab.B200Array.empty(rt, (max(a.size, 1),), a.dtype)
    _lib.call("dab_sort", rt.ctx, dab.dab_dtype(a.dtype), C.c_void_p(src.ptr), C.c_void_p(out.ptr), C.c_void_p(tmp.ptr), a.size)
    got = out.to_numpy()
    if not inplace:
        assert np.array_equal(src.to_numpy().view(np.uint8), a.view(np.uint8))   # the input is never written
        src.free()
    out.free()
    tmp.free()
    return got


@pytest.mark.parametrize("T", [np.int64, np.float64, np.int32, np.float32])
def test_sort_kernel_sizes_and_patterns(dab, rt1, T):
    rng = np.random.default_rng(51)
    for n in (0, 1, 2, 31, 1024, 1025, 4096, 4097, 100003, (1 << 20) + 17):
        for kind in ("full", "small"):
            a = _data(T, n, rng, kind)
            for inplace in (False, True):
                got = _sort_chunk(dab, rt1, a, inplace)
                assert got.dtype == a.dtype and np.array_equal(got, orc.jl_sort(a))
    n = 50000
    for a in (np.zeros(n, dtype=T), np.arange(n).astype(T), np.arange(n)[::-1].astype(T), np.repeat(np.arange(50), n // 50).astype(T)):
        assert np.array_equal(_sort_chunk(dab, rt1, np.ascontiguousarray(a)), np.sort(a))


def test_sort_kernel_float_order(dab, rt1):
    """isless order: -Inf < negatives < -0.0 < +0.0 < positives < Inf < NaN (either sign), bit patterns preserved."""
    rng = np.random.default_rng(52)
    for T, U in ((np.float32, np.uint32), (np.float64, np.uint64)):
        a = _data(T, 20000, rng)
        a[rng.integers(0, a.size, 300)] = 0.0
        a[rng.integers(0, a.size, 300)] = -0.0
        a[rng.integers(0, a.size, 50)] = np.inf
        a[rng.integers(0, a.size, 50)] = -np.inf
        a[rng.integers(0, a.size, 100)] = np.nan
        a[rng.integers(0, a.size, 100)] = -np.nan
        got = _sort_chunk(dab, rt1, a)
        want = orc.jl_sort(a)
        k = int((~np.isnan(a)).sum())
        assert np.array_equal(got[:k].view(U), want[:k].view(U))                 # incl. the sign of the zeros
        assert np.all(np.isnan(got[k:])) and got.size == a.size
        assert sorted(got[k:].view(U).tolist()) == sorted(a[np.isnan(a)].view(U).tolist())   # NaN payloads and signs survive


def test_sorted_split_matches_reference_scan(dab, rt1):
    from darray_b200 import _lib
    rng = np.random.default_rng(53)
    for T in (np.int64, np.float64, np.float32, np.int32):
        a = orc.jl_sort(_data(T, 30011, rng, "small"))
        if np.dtype(T).kind == "f":
            a[-7:] = np.nan
            a[1000:1010] = 0.0
            a[990:1000] = -0.0
            a = orc.jl_sort(a)
        lo, hi = (np.iinfo(T).min, np.iinfo(T).max) if np.dtype(T).kind == "i" else (-np.inf, np.inf)
        inner = np.sort(rng.choice(a[~np.isnan(a.astype(np.float64))], 6))
        bounds = np.concatenate([[lo], inner, [hi]]).astype(T)
        if np.dtype(T).kind == "f":
            bounds[2] = -0.0
        bounds = np.sort(bounds)
        want = orc.sort_split_points(a, bounds)
        dev = dab.B200Array.from_numpy(rt1, a)
        cnt = (C.c_ulonglong * (len(bounds) - 1))()
        b1 = np.ascontiguousarray(bounds[1:])
        _lib.call("dab_sorted_split", rt1.ctx, dab.dab_dtype(T), C.c_void_p(dev.ptr), a.size, C.c_void_p(b1.ctypes.data), len(b1), cnt)
        ends, prev = [], 0
        for c in cnt:
            prev = max(prev, int(c))
            ends.append(prev)
        assert ends == want
        dev.free()


@pytest.mark.parametrize("T", [np.int64, np.float64])
@pytest.mark.parametrize("i", range(0, 7))
def test_darray_sort_reference_test(dab, rt8, T, i):
    """test/darray.jl:1015-1025 with the oracle checking layout and boundaries as well."""
    rng = np.random.default_rng(200 + i)
    n = 10 ** i
    a = _data(T, n, rng)
    d = dab.distribute(a)
    od = orc.distribute(a, nworkers=8)
    assert list(d.layout.pids) == od.pids
    for sample in (True, False, (a.min(), a.max()), _data(T, min(n, 512), rng)):
        d2, b = dab.sort_with_boundaries(d, sample=sample)
        o2, ob = orc.darray_sort(od, sample)
        assert np.array_equal(b, ob)
        assert len(d2) == n and np.array_equal(dab.to_array(d2), np.sort(a))
        assert list(d2.layout.pids) == o2.pids and list(d2.layout.indices) == o2.indices and d2.layout.cuts[0] == o2.cuts[0]
        for pid, ch in d2.chunks.items():
            assert np.array_equal(ch.to_numpy(), o2.chunks[o2.pids.index(pid)])
        d2.close()
    assert np.array_equal(dab.to_array(dab.sort(d)), np.sort(a))                  # the public spelling


def test_darray_sort_other_types_and_errors(dab, rt8):
    rng = np.random.default_rng(61)
    for T in (np.float32, np.int32):
        a = _data(T, 77777, rng, "small")
        d = dab.distribute(a, procs=[1, 2, 3], dist=[3])
        o2, _ = orc.darray_sort(orc.distribute(a, procs=[1, 2, 3], dist=[3]), True)
        d2 = dab.sort(d)
        assert np.array_equal(dab.to_array(d2), np.sort(a)) and list(d2.layout.indices) == o2.indices
    # a heavily skewed vector: some workers receive nothing and drop out of the result (src/sort.jl:163-168)
    a = np.concatenate([np.zeros(5000, dtype=np.int64), np.arange(8, dtype=np.int64)])
    d = dab.distribute(a)
    od = orc.distribute(a, nworkers=8)
    d2 = dab.sort(d)
    o2, _ = orc.darray_sort(od, True)
    assert list(d2.layout.pids) == o2.pids and len(o2.pids) < 8 and np.array_equal(dab.to_array(d2), np.sort(a))
    with pytest.raises(dab.ArgumentError):
        dab.sort(d, rev=True)                                                     # only alg, by, sample (src/sort.jl:112-114)
    with pytest.raises(dab.ArgumentError):
        dab.sort(d, sample="yes")
    with pytest.raises(TypeError):
        dab.sort(d, by=lambda x: x if x > 0 else -x)                              # a key function must be traceable (no data-dependent branches)
    with pytest.raises(dab.ArgumentError):
        dab.sort(dab.distribute(rng.random(100)), sample=(-np.inf, 1.0))
