"""GPU parity tests for the rows SURVEY.md section 8(f) marks "next" that are built so far: fused general mapreduce (arbitrary
closure, several arguments), Level-1 linear algebra (reference src/linalg.jl:24-59, tests test/darray.jl:258-266, 930-965),
`==` (src/darray.jl:403-414) and `mean` (ext/StatisticsExt.jl:6, tests test/darray.jl:336-348)."""
import numpy as np
import pytest

from oracle import darray_oracle as orc

pytestmark = pytest.mark.gpu
F32 = np.float32


def test_general_mapreduce_is_fused_and_exact_for_ints(dab, rt8):
    """reference test/darray.jl:286-294 again, now through the single-pass NVRTC map+reduce kernel."""
    rng = np.random.default_rng(17)
    A = rng.integers(1, 6, 100003).astype(np.int64)
    DA = dab.distribute(A)
    l0 = rt8.launches()
    assert dab.mapreduce(lambda x: x ** 2 + 2 * x - 1, "+", DA) == (A * A + 2 * A - 1).sum()
    assert rt8.launches() - l0 == 2 * len(DA.chunks)            # partial + final kernel per chunk, no temporary f.(d) pass
    assert dab.mapreduce(lambda x: 2 * x, "max", DA) == (2 * A).max()
    assert dab.mapreduce(lambda x: x % 3, "min", DA) == 0
    B = rng.integers(1, 3, 40).astype(np.int64)
    assert dab.mapreduce(lambda x: 2 * x, "*", dab.distribute(B)) == np.prod(2 * B)
    x = orc.rand_u01(3, 0, (1 << 20) + 5)
    d = dab.distribute(x)
    got = dab.mapreduce(lambda v: dab.sqrt(v) * v + 1, "+", d)
    want = (np.sqrt(x.astype(np.float64)) * x + 1).sum()
    assert got.dtype == np.float32 and abs(float(got) - want) <= 2e-6 * want
    assert dab.count(d, lambda v: (v > 0.25) & (v < 0.5)) == int(((x > 0.25) & (x < 0.5)).sum())
    assert dab.all(d, lambda v: v * v <= v) and not dab.any(d, lambda v: v + 1 < 1)
    assert dab.mapreduce(lambda v: v > 0.5, "+", d) == int((x > 0.5).sum())   # sum of Bools -> Int64 count


def test_dot_norm_axpy_rmul(dab, rt8):
    """reference test/darray.jl:930-965 and :258-266."""
    rng = np.random.default_rng(23)
    for dtype, tol in ((np.float64, 1.5e-8), (np.float32, 3e-6)):
        n = 200003
        xa, ya = rng.standard_normal(n).astype(dtype), rng.standard_normal(n).astype(dtype)
        x, y = dab.distribute(xa), dab.distribute(ya)
        d = dab.dot(x, y)
        ref = np.dot(xa.astype(np.float64), ya.astype(np.float64))
        assert d.dtype == dtype and abs(float(d) - ref) <= tol * np.sqrt(n)
        x64 = xa.astype(np.float64)
        assert abs(float(dab.norm(x)) - np.linalg.norm(x64)) < tol * np.linalg.norm(x64)
        assert abs(float(dab.norm(x, 2)) - np.linalg.norm(x64, 2)) < tol * np.linalg.norm(x64)
        assert abs(float(dab.norm(x, 1)) - np.linalg.norm(x64, 1)) < tol * np.linalg.norm(x64, 1)
        assert float(dab.norm(x, np.inf)) == np.abs(xa).max()
        yc = dab.distribute(ya)
        dab.axpy_(2.0, x, yc)
        assert np.array_equal(dab.to_array(yc), dtype(2.0) * xa + ya)           # unfused: bit-exact vs NumPy
        with pytest.raises(dab.DimensionMismatch):
            dab.axpy_(2.0, x, dab.distribute(np.zeros(n + 1, dtype=dtype)))
        dab.rmul_(x, 2)
        assert np.array_equal(dab.to_array(x), xa * dtype(2))
    # dot of DArrays with different layouts (makelocal halo fetch inside, reference src/linalg.jl:41)
    A = rng.standard_normal((64, 48))
    a, b = dab.distribute(A, dist=(8, 1)), dab.distribute(A, dist=(2, 4))
    assert np.isclose(float(dab.dot(a, b)), (A * A).sum(), rtol=1e-12)


def test_isequal(dab, rt8):
    """reference src/darray.jl:403-414 and test/darray.jl:20-35, 264."""
    rng = np.random.default_rng(29)
    A = rng.standard_normal((50, 30))
    DA = dab.distribute(A)
    assert dab.isequal(DA, A) and dab.isequal(DA, dab.distribute(A, dist=(1, 8)))
    B = A.copy()
    B[49, 29] += 1e-9
    assert not dab.isequal(DA, B) and not dab.isequal(DA, A[:, :29]) and not dab.isequal(DA, dab.distribute(B))
    N = A.copy()
    N[3, 3] = np.nan
    assert not dab.isequal(dab.distribute(N), N)          # NaN != NaN, as in Julia


@pytest.mark.parametrize("dms", [1, 2, 3, (1, 2), (1, 3), (2, 3), (1, 2, 3)])
def test_reference_mean_dims(dab, rt8, dms):
    """reference test/darray.jl:336-348."""
    rng = np.random.default_rng(31)
    A = rng.standard_normal((20, 20, 20))
    DA = dab.distribute(A)
    ax = tuple(d - 1 for d in ((dms,) if isinstance(dms, int) else dms))
    assert np.allclose(dab.to_array(dab.mean(DA, dims=dms)), A.mean(axis=ax, keepdims=True), rtol=1.5e-8)
    assert np.isclose(float(dab.mean(DA)), A.mean(), rtol=1e-12)
    Af = A.astype(F32)
    m = dab.mean(dab.distribute(Af), dims=dms)
    assert m.dtype == np.float32 and np.allclose(dab.to_array(m), Af.astype(np.float64).mean(axis=ax, keepdims=True), rtol=1e-5, atol=1e-6)


def test_full_size_dot_and_general_mapreduce_rate(dab, rt1):
    """2^29-element dot and a general closure at full size: single pass each; values against the exact integer identities of the
    counter-based input (x_i = k_i 2^-24): sum(x .* 1) and sum(2x+1)."""
    n = 1 << 29
    x = dab.drand((n,), dtype=F32, seed=5)
    ones = dab.dones((n,), dtype=F32)
    from oracle import core as ocore

    exact = 0
    blk = 1 << 24
    for s in range(0, n, blk):
        exact += ocore.rand_ksum(5, s, blk)
    exact *= 2.0 ** -24
    d = float(dab.dot(x, ones))
    assert abs(d - exact) <= 1e-6 * exact
    g = float(dab.mapreduce(lambda v: 2 * v + 1, "+", x))
    assert abs(g - (2 * exact + n)) <= 1e-6 * (2 * exact + n)


def test_map_localparts_operators(dab, rt8):
    """reference src/mapreduce.jl:134-189: + - div mod rem & | xor between DArrays / Arrays of ONE element type; the result keeps
    the first DArray's layout; a differently cut second DArray is brought to it (samedist)."""
    rng = np.random.default_rng(37)
    A = rng.integers(-40, 40, (30, 44)).astype(np.int64)
    B = rng.integers(1, 9, (30, 44)).astype(np.int64)
    a = dab.distribute(A, dist=(2, 4))
    b = dab.distribute(B, dist=(8, 1))                       # different cuts -> samedist
    for got, want in ((a + b, A + B), (a - b, A - B), (a & b, A & B), (a | b, A | B), (a ^ b, A ^ B), (a // b, np.trunc(A / B).astype(np.int64)),
                      (a % b, np.fmod(A, B)), (-a, -A), (a + B, A + B), (A - b, A - B)):
        assert isinstance(got, dab.DArray) and np.array_equal(dab.to_array(got), want)
    s = a + b
    assert s.layout.grid == (2, 4) and s.indices == a.indices and s.layout.pids == a.layout.pids     # layout of the first DArray
    assert (A - b).indices == b.indices
    F = rng.standard_normal((30, 44))
    f = dab.distribute(F)
    assert np.array_equal(dab.to_array(f + f), F + F) and np.array_equal(dab.to_array(f - F), F - F)
    with pytest.raises(dab.DimensionMismatch):
        a + dab.distribute(A[:, :40])
    with pytest.raises(TypeError):
        a + f                                              # Int64 + Float64 DArrays: no such method in the reference
