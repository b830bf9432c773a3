#!/usr/bin/env python
"""Single-GPU rates of the widened rows (fused NVRTC map+reduce, Level-1 ops): algorithmic GB/s, CUDA events, 2^30 Float32."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darray_b200 as dab  # noqa: E402

rt = dab.init(use_dist=False)
n = 1 << 30
x = dab.drand((n,), dtype=np.float32, seed=1)
y = dab.drand((n,), dtype=np.float32, seed=2)


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


for name, fn, b in [("dot(x, y)                      8 B/elem", lambda: dab.dot(x, y), 8 * n),
                    ("mapreduce(v -> 2v+1, +, x)     4 B/elem", lambda: dab.mapreduce(lambda v: 2 * v + 1, "+", x), 4 * n),
                    ("norm(x)                        4 B/elem", lambda: dab.norm(x), 4 * n),
                    ("x == y (all(x .== y))          8 B/elem", lambda: dab.isequal(x, y), 8 * n),
                    ("count(v -> 0.25 < v < 0.5, x)  4 B/elem", lambda: dab.count(x, lambda v: (v > 0.25) & (v < 0.5)), 4 * n),
                    ("axpy!(2, x, y)                12 B/elem", lambda: dab.axpy_(2.0, x, y), 12 * n),
                    ("rmul!(x, 1.0000001)            8 B/elem", lambda: dab.rmul_(x, 1.0000001), 8 * n)]:
    ms = timed(fn)
    print(f"{name:44s} {ms:8.4f} ms {b / ms / 1e6:9.1f} GB/s", flush=True)
