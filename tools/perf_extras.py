#!/usr/bin/env python
"""Single-GPU throughput table for the secondary kernels of the hot path (not the bench.py headline): mapreducedim on the C4
chunk shape, local box copies, fill!/rand!, binary and NVRTC-fused broadcasts, other element types.  Prints one line per case:
algorithmic GB/s (CUDA events on the ctx stream, inputs >> L2)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import darray_b200 as dab  # noqa: E402
from darray_b200 import _lib  # noqa: E402

F32 = np.float32
rt = dab.init(workers_per_rank=1, use_dist=False)
out = {}


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = rt.event(), rt.event()
    rt.sync()
    rt.record(e0)
    for _ in range(reps):
        fn()
    rt.record(e1)
    ms = rt.elapsed_ms(e0, e1) / reps
    rt.event_destroy(e0)
    rt.event_destroy(e1)
    return ms


def report(name, nbytes, ms):
    out[name] = round(nbytes / ms / 1e6, 1)
    print(f"{name:58s} {ms:9.4f} ms  {nbytes / ms / 1e6:9.1f} GB/s", flush=True)


# ---- C4 chunk: 32768 x 16384 Float32 (2 GiB), sum(dims=1) [contiguous runs] and sum(dims=2) [strided]
R, Cc = 32768, 16384
A = dab.drand((R, Cc), dtype=F32, seed=1)
ch = A.chunks[1]
o1 = dab.B200Array.empty(rt, (Cc,), F32)
o2 = dab.B200Array.empty(rt, (R,), F32)
call = lambda inner, red, outer, o, op=_lib.SUM: _lib.call("dab_reducedim", rt.ctx, _lib.F32, op, _lib.MAP_ID, C.c_void_p(ch.ptr), inner, red, outer, C.c_void_p(o.ptr), 0)
report("reducedim C4 chunk sum(dims=1)  (1, 32768, 16384)", 4 * R * Cc, timed(lambda: call(1, R, Cc, o1)))
report("reducedim C4 chunk sum(dims=2)  (32768, 16384, 1)", 4 * R * Cc, timed(lambda: call(R, Cc, 1, o2)))
report("reducedim C4 chunk max(dims=1)", 4 * R * Cc, timed(lambda: call(1, R, Cc, o1, _lib.MAX)))
report("reducedim whole chunk as one run (1, 2^29, 1)", 4 * R * Cc, timed(lambda: call(1, R * Cc, 1, o1)))
big_out = dab.B200Array.empty(rt, (R * Cc // 64,), F32)
report("reducedim short runs (1, 64, 2^23)", 4 * R * Cc, timed(lambda: call(1, 64, R * Cc // 64, big_out)))
report("reducedim mid runs (1, 1000, 536870)", 4 * 1000 * 536870, timed(lambda: call(1, 1000, 536870, big_out)))
report("reducedim strided (64, 128, 65536)", 4 * R * Cc, timed(lambda: call(64, 128, 65536, big_out)))
# through the DArray API (phase 1 + between-phase + allocation of R)
report("DArray sum(A, dims=1) API, 1 worker", 4 * R * Cc, timed(lambda: dab.sum(A, dims=1).close(), reps=5))

# ---- local box copies (HBM -> HBM: 2 x bytes of traffic; report bytes MOVED, i.e. 4 B/elem)
n = 1 << 28
src = dab.B200Array.empty(rt, (n + 64,), F32)
dst = dab.B200Array.empty(rt, (n + 64,), F32)
cp = lambda so, do, ext, ss, ds, es=4: _lib.call("dab_copy_box", rt.ctx, es, C.c_void_p(dst.ptr), _lib.sz4(ds), _lib.sz4(list(do) + [0] * (4 - len(do))), C.c_void_p(src.ptr), _lib.sz4(ss),
                                                  _lib.sz4(list(so) + [0] * (4 - len(so))), _lib.sz4(ext))
report("copy_box contiguous aligned 1 GiB (moved bytes)", 4 * n, timed(lambda: cp([0], [0], [n], [n + 64], [n + 64])))
report("copy_box contiguous src+4B misaligned", 4 * n, timed(lambda: cp([1], [0], [n], [n + 64], [n + 64])))
report("copy_box 2-D strided 8192x8192 of 16384 rows", 4 * 8192 * 8192, timed(lambda: cp([1024, 0], [0, 0], [8192, 8192], [16384, 8192 * 2], [8192, 8192])))
report("copy_box 2-D strided, rows of 100 floats", 4 * 100 * 2000000, timed(lambda: cp([3, 0], [0, 0], [100, 2000000], [128, 2000000], [100, 2000000])))

# ---- fill! / rand! (write-only 4 B/elem)
n = 1 << 30
x = dab.drand((n,), dtype=F32, seed=2)
y = dab.similar(x)
z = dab.similar(x)
one = np.asarray(1.0, dtype=F32)
report("fill! 2^30 f32", 4 * n, timed(lambda: _lib.call("dab_fill", rt.ctx, _lib.F32, C.c_void_p(y.chunks[1].ptr), n, C.c_void_p(one.ctypes.data))))
report("rand! 2^30 f32", 4 * n, timed(lambda: _lib.call("dab_rand_u01", rt.ctx, _lib.F32, C.c_void_p(y.chunks[1].ptr), n, 7, 0)))
# ---- broadcasts
report("broadcast z .= x .+ y (binary, 12 B/elem)", 12 * n, timed(lambda: dab.broadcast_into(z, lambda u, v: u + v, x, y)))
report("broadcast z .= x .* 2f0 (binary_scalar, 8 B/elem)", 8 * n, timed(lambda: dab.broadcast_into(z, lambda u: u * F32(2), x)))
report("broadcast z .= abs.(x) (unary, 8 B/elem)", 8 * n, timed(lambda: dab.broadcast_into(z, lambda u: abs(u), x)))
report("broadcast z .= x .- y .* x (NVRTC fused, 12 B/elem)", 12 * n, timed(lambda: dab.broadcast_into(z, lambda u, v: u - v * u, x, y)))
report("broadcast z .= sqrt.(abs2.(x) .+ abs2.(y)) (NVRTC, 12 B/elem)", 12 * n, timed(lambda: dab.broadcast_into(z, lambda u, v: dab.sqrt(dab.abs2(u) + dab.abs2(v)), x, y)))
report("sum(abs2, x) f32", 4 * n, timed(lambda: dab.sum(x, dab.abs2)))
# extruded operands (reference test/darray.jl:885-898): a .- m with a 1 x n row m, a .* v with a column v; 8 B/elem
M2 = dab.drand((16384, 16384), dtype=F32, seed=9)
Z2 = dab.similar(M2)
mrow = dab.drand((1, 16384), dtype=F32, seed=10)
vcol = dab.drand((16384, 1), dtype=F32, seed=12)
report("broadcast Z .= A .- m   (m 1 x n, extruded; NVRTC rows)", 8 * 16384 * 16384, timed(lambda: dab.broadcast_into(Z2, lambda u, v: u - v, M2, mrow)))
report("broadcast Z .= A .* v   (v n x 1, extruded; NVRTC rows)", 8 * 16384 * 16384, timed(lambda: dab.broadcast_into(Z2, lambda u, v: u * v, M2, vcol)))
M2.close(); Z2.close()
report("count(x .> 0.5) f32", 4 * n, timed(lambda: dab.count(x, lambda v: v > 0.5)))
report("extrema(x) f32 (2 passes)", 8 * n, timed(lambda: dab.extrema(x)))
x.close(); y.close(); z.close()
n = 1 << 29
xd = dab.drand((n,), dtype=np.float64, seed=3)
report("sum f64 2^29", 8 * n, timed(lambda: dab.sum(xd)))
report("maximum f64 2^29", 8 * n, timed(lambda: dab.maximum(xd)))
yd = dab.similar(xd)
report("broadcast y .= 1.5 .* x .+ 0.25 f64 (16 B/elem)", 16 * n, timed(lambda: dab.broadcast_into(yd, lambda v: 1.5 * v + 0.25, xd)))
print(json.dumps(out))
