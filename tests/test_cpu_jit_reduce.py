"""CPU-only: the source generated for the fused map+reduce kernels (dab_mapreduce_expr) compiles with NVRTC for sm_100a for every
(op, value type) combination the host runtime can request, and unsupported combinations are refused, not silently served."""
import ctypes as C


def _check(dab, f, tags, op, arrays=None):
    from darray_b200 import _lib
    from darray_b200._broadcast import codegen, trace

    code = {"f32": 0, "f64": 1, "i32": 2, "i64": 3, "bool": 4}
    e = trace(f, tags)
    n = len(tags)
    dts = (C.c_int32 * n)(*[code[t] for t in tags])
    arr = (C.c_int32 * n)(*(arrays or [1] * n))
    sz = C.c_size_t()
    L = _lib.lib()
    st = L.dab_jit_compile_check_reduce(codegen(e).encode(), code[e.jt], op, n, dts, arr, C.byref(sz))
    return st, sz.value, (L.dab_last_error(None) or b"").decode()[:300]


def test_fused_mapreduce_codegen_compiles(dab):
    from darray_b200 import _lib, sqrt

    cases = [(lambda a, b: a * b, ["f32", "f32"], _lib.SUM), (lambda a, b: a * b, ["f64", "f64"], _lib.SUM),
             (lambda x: x ** 2 + 2 * x - 1, ["i64"], _lib.SUM), (lambda x: 2 * x, ["i32"], _lib.PROD), (lambda x: sqrt(x) * x + 1, ["f32"], _lib.SUM),
             (lambda a, b: a == b, ["f64", "f64"], _lib.ALL), (lambda v: (v > 0.25) & (v < 0.5), ["f32"], _lib.COUNT),
             (lambda v: v + 1 < 1, ["f32"], _lib.ANY), (lambda v: v > 0.5, ["f32"], _lib.SUM), (lambda x: 2 * x, ["i64"], _lib.MAX),
             (lambda x, s: x * s, ["f32", "f32"], _lib.MIN), (lambda x: -x, ["f64"], _lib.MAX), (lambda x: x % 3, ["i32"], _lib.MIN)]
    for f, tags, op in cases:
        st, size, err = _check(dab, f, tags, op, arrays=[1] + [0] * (len(tags) - 1) if len(tags) == 2 and tags[1] == "f32" and op == _lib.MIN else None)
        assert st == 0 and size > 2000, (tags, op, err)


def test_fused_mapreduce_refuses_unsupported(dab):
    from darray_b200 import _lib

    st, _, err = _check(dab, lambda v: v > 0.5, ["f32"], _lib.MAX)        # max of Bools: not served
    assert st == _lib.ERR_UNSUPPORTED
    st, _, err = _check(dab, lambda v: v * 2, ["f32"], _lib.ALL)          # all() of non-Bool values
    assert st == _lib.ERR_UNSUPPORTED
