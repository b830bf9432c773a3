#!/usr/bin/env python
"""Single-GPU rates of the Level-2 widening (K9 dab_gemv, K10 dab_transpose_box, Diagonal scaling): algorithmic GB/s with CUDA
events.  Shapes: the C4 chunk (32768 x 16384 Float32 = 2 GiB), a tall and a wide variant, and Float64."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darray_b200 as dab  # noqa: E402
from darray_b200 import _lib  # noqa: E402

rt = dab.init(use_dist=False)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    e0, e1 = rt.event(), rt.event()
    rt.sync()
    rt.record(e0)
    for _ in range(reps):
        fn()
    rt.record(e1)
    return rt.elapsed_ms(e0, e1) / reps


for dtype, shapes in ((np.float32, [(32768, 16384), (1 << 22, 128), (128, 1 << 22), (1 << 29, 1), (1, 1 << 29), (32767, 16385)]),
                      (np.float64, [(32768, 8192)])):
    es = np.dtype(dtype).itemsize
    for (m, n) in shapes:
        A = dab.drand((m, n), dtype=dtype, seed=3, procs=[1], dist=(1, 1))
        ch = A.chunks[1]
        for trans in (0, 1):
            x = dab.B200Array.empty(rt, (m if trans else n,), dtype)
            r = dab.B200Array.empty(rt, (n if trans else m,), dtype)
            one = np.ones((), dtype=dtype)
            _lib.call("dab_fill", rt.ctx, dab.dab_dtype(dtype), C.c_void_p(x.ptr), x.size, C.c_void_p(one.ctypes.data))
            fn = lambda: _lib.call("dab_gemv", rt.ctx, dab.dab_dtype(dtype), trans, C.c_void_p(ch.ptr), m, n, C.c_void_p(x.ptr),  # noqa: E731
                                   C.c_void_p(r.ptr))
            ms = timed(fn)
            print(f"gemv {'T' if trans else 'N'} {np.dtype(dtype).name} {m:>10d} x {n:<10d} {ms:8.4f} ms {m * n * es / ms / 1e6:9.1f} GB/s",
                  flush=True)
            x.free()
            r.free()
        if m * n <= (1 << 29) and min(m, n) > 1:
            ms = timed(lambda: dab.lmul_diag(np.ones(m, dtype=dtype), A), reps=5)
            print(f"lmul!(Diagonal, A) {np.dtype(dtype).name} {m} x {n}   {ms:8.4f} ms {2 * m * n * es / ms / 1e6:9.1f} GB/s (includes H2D of the diagonal)",
                  flush=True)
        A.close()

for dtype, (m, n) in ((np.float32, (32768, 16384)), (np.float32, (16384, 32768)), (np.float64, (16384, 16384)), (np.float32, (32767, 16385))):
    es = np.dtype(dtype).itemsize
    A = dab.drand((m, n), dtype=dtype, seed=4, procs=[1], dist=(1, 1))
    out = dab.B200Array.empty(rt, (n, m), dtype)
    src = A.chunks[1]
    fn = lambda: _lib.call("dab_transpose_box", rt.ctx, es, C.c_void_p(out.ptr), n, C.c_void_p(src.ptr), m, m, n)  # noqa: E731
    ms = timed(fn)
    print(f"transpose {np.dtype(dtype).name} {m} x {n}  {ms:8.4f} ms {2 * m * n * es / ms / 1e6:9.1f} GB/s (read + write)", flush=True)
    out.free()
    A.close()
dab.d_closeall()
rt.shutdown()
