#!/usr/bin/env python
"""A*x (K9, trans = 0) on Float32 / Float64 chunks whose columns are NOT 16-byte aligned: the phase-class kernel (default) against the
round-1 pair (dab_set_option gemv_phase=0: single-wave aligned kernel / unit-wise loads), aligned shapes next to misaligned ones; then
A'*x (trans = 1): unit-wise kernel (gemv_phase=0) against the phase-class kernel over the CTA waves (gemv_t_waves) and the columns a
thread carries (gemv_t_cols).  CUDA events, 10 reps, algorithmic GB/s."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darray_b200 as dab  # noqa: E402
from darray_b200 import _lib  # noqa: E402

rt = dab.init(use_dist=False)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = rt.event(), rt.event()
    rt.sync()
    rt.record(e0)
    for _ in range(reps):
        fn()
    rt.record(e1)
    return rt.elapsed_ms(e0, e1) / reps


for dtype, shapes in ((np.float32, [(32768, 16384), (32767, 16385), (32769, 16384), (36001, 14001), (32770, 16384), (1 << 20, 509), (4099, 131071)]),
                      (np.float64, [(32768, 8192), (32767, 8193)])):
    es = np.dtype(dtype).itemsize
    for (m, n) in shapes:
        A = dab.drand((m, n), dtype=dtype, seed=3, procs=[1], dist=(1, 1))
        ch = A.chunks[1]
        x = dab.B200Array.empty(rt, (n,), dtype)
        r = dab.B200Array.empty(rt, (m,), dtype)
        one = np.ones((), dtype=dtype)
        _lib.call("dab_fill", rt.ctx, dab.dab_dtype(dtype), C.c_void_p(x.ptr), x.size, C.c_void_p(one.ctypes.data))
        fn = lambda: _lib.call("dab_gemv", rt.ctx, dab.dab_dtype(dtype), 0, C.c_void_p(ch.ptr), m, n, C.c_void_p(x.ptr), C.c_void_p(r.ptr))  # noqa: E731
        res = {}
        for phase in (1, 0):
            rt.set_option("gemv_phase", phase)
            ms = timed(fn)
            res[phase] = (ms, r.to_numpy())
            print(f"gemv N {np.dtype(dtype).name} {m:>8d} x {n:<8d} gemv_phase={phase}: {ms:8.4f} ms {m * n * es / ms / 1e6:8.1f} GB/s", flush=True)
        rt.set_option("gemv_phase", 1)
        same = np.all(np.abs(res[1][1] - res[0][1]) <= np.spacing(np.abs(res[0][1])))
        print(f"    results agree to 1 ulp: {bool(same)}", flush=True)
        x.free()
        r.free()
        xt = dab.B200Array.empty(rt, (m,), dtype)
        rr = dab.B200Array.empty(rt, (n,), dtype)
        _lib.call("dab_fill", rt.ctx, dab.dab_dtype(dtype), C.c_void_p(xt.ptr), xt.size, C.c_void_p(one.ctypes.data))
        fnt = lambda: _lib.call("dab_gemv", rt.ctx, dab.dab_dtype(dtype), 1, C.c_void_p(ch.ptr), m, n, C.c_void_p(xt.ptr), C.c_void_p(rr.ptr))  # noqa: E731
        ref = None
        for phase, waves, cols in ((0, 4, 8), (1, 1, 4), (1, 4, 4), (1, 1, 8), (1, 4, 8)):
            rt.set_option("gemv_phase", phase)
            rt.set_option("gemv_t_waves", waves)
            rt.set_option("gemv_t_cols", cols)
            ms = timed(fnt)
            out = rr.to_numpy()
            ref = out if ref is None else ref
            ok = bool(np.all(np.abs(out - ref) <= np.spacing(np.abs(ref))))
            print(f"gemv T {np.dtype(dtype).name} {m:>8d} x {n:<8d} gemv_phase={phase} waves={waves} cols={cols}: {ms:8.4f} ms {m * n * es / ms / 1e6:8.1f} GB/s"
                  f"  (= first to 1 ulp: {ok})", flush=True)
        rt.set_option("gemv_phase", 1)
        rt.set_option("gemv_t_waves", 4)
        rt.set_option("gemv_t_cols", 8)
        xt.free()
        rr.free()
        A.close()
dab.d_closeall()
rt.shutdown()
