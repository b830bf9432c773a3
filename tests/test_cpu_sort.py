"""CPU tests of the samplesort restatement (reference src/sort.jl; its own test is test/darray.jl:1015-1025:
``sort(Array(d)) == Array(sort(d; sample=s))`` for 10^0..10^6 elements, T in (Int, Float64), four kinds of ``sample``)."""
import numpy as np
import pytest

from oracle import darray_oracle as orc


def _data(T, n, rng):
    if np.dtype(T).kind == "i":
        return rng.integers(np.iinfo(T).min, np.iinfo(T).max, n, dtype=T)      # rand(Int, n): the full range
    return rng.random(n).astype(T)


@pytest.mark.parametrize("T", [np.int64, np.float64])
@pytest.mark.parametrize("i", range(0, 6))
def test_oracle_sort_reference_test(T, i):
    rng = np.random.default_rng(100 + i)
    n = 10 ** i
    a = _data(T, n, rng)
    for nw in (1, 2, 8):
        d = orc.distribute(a, nworkers=nw)
        for sample in (True, False, (a.min(), a.max()), _data(T, min(n, 512), rng)):
            d2, b = orc.darray_sort(d, sample)
            a2 = orc.to_array(d2)
            assert len(a2) == n and np.array_equal(np.sort(a), a2)
            assert len(b) == len(d.pids) + 1 and np.all(b[1:-1][:-1] <= b[1:-1][1:])
            assert d2.pids == [d.pids[k] for k in range(len(d.pids)) if k in [d.pids.index(p) for p in d2.pids]]
            assert all(len(c) > 0 for c in d2.chunks) and sum(len(c) for c in d2.chunks) == n
            assert d2.cuts[0][0] == 1 and d2.cuts[0][-1] == n + 1


def test_oracle_sort_details():
    # the sample picks sorted[1:div(llp,ss):llp]: more than 512 keys when 512 < llp < 1024 (step 1)
    assert list(orc.sort_sample_indices(1000)) == list(range(1000))
    assert list(orc.sort_sample_indices(2048)) == list(range(0, 2048, 4))
    assert len(orc.sort_sample_indices(1025)) == 513
    with pytest.raises(ZeroDivisionError):
        orc.sort_sample_indices(0)
    # boundaries: first sample replaced by typemin, typemax appended
    b = orc.sort_boundaries_from_samples(np.array([5, 1, 9, 3, 7, 2], dtype=np.int64), 3, np.int64)
    assert list(b) == [np.iinfo(np.int64).min, 3, 7, np.iinfo(np.int64).max]
    # split scan: piece i = leading run of elements NOT > boundaries[i+1]
    s = np.array([1, 2, 3, 3, 4, 8, 9], dtype=np.int64)
    assert orc.sort_split_points(s, b) == [4, 5, 7]
    # isless order: -0.0 before +0.0, NaN last, and a NaN never exceeds a boundary (it stays in the piece being scanned)
    v = np.array([0.0, np.nan, -0.0, 1.0, -np.inf], dtype=np.float64)
    sv = orc.jl_sort(v)
    assert np.signbit(sv[1]) and not np.signbit(sv[2]) and sv[0] == -np.inf and np.isnan(sv[-1])
    bf = np.array([-np.inf, 0.0, np.inf])
    assert orc.sort_split_points(sv, bf) == [3, 5]
    # uniform sample with a full-range Int64 min/max: abs(ub - lb) wraps like the reference's machine arithmetic
    lo, hi = np.int64(-9 * 10 ** 18), np.int64(9 * 10 ** 18)
    u = orc.sort_uniform_sample(lo, hi, 4, np.int64)
    wrapped = (int(hi) - int(lo) + 2 ** 63) % 2 ** 64 - 2 ** 63
    assert u[0] == lo and u[1] == np.int64(np.rint(float(lo) + abs(wrapped) / 4))
    with pytest.raises(ValueError):
        orc.sort_uniform_sample(-np.inf, 1.0, 2, np.float64)


def test_host_boundary_logic_matches_oracle():
    from darray_b200 import _sort
    rng = np.random.default_rng(3)
    for T in (np.int64, np.float64, np.float32, np.int32):
        dt = np.dtype(T)
        s = _data(T, 700, rng)
        for nparts in (1, 2, 3, 8):
            assert np.array_equal(_sort.boundaries_from_samples(s, nparts, dt), orc.sort_boundaries_from_samples(s, nparts, dt))
            assert np.array_equal(_sort.uniform_sample(s.min(), s.max(), nparts, dt), orc.sort_uniform_sample(s.min(), s.max(), nparts, dt))
