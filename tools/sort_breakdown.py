#!/usr/bin/env python
"""One dab_sort of 2^27 Int64 and one of 2^28 Float32 after a warm-up: run under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sort_launches.csv python tools/sort_breakdown.py
to get the per-kernel times of the radix sort (histogram / count / scan / scatter per pass)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darray_b200 as dab  # noqa: E402
from darray_b200 import _lib  # noqa: E402

rt = dab.init(use_dist=False)
rng = np.random.default_rng(1)
for T, n in ((np.int64, 1 << 27), (np.float32, 1 << 28)):
    a = rng.integers(np.iinfo(T).min, np.iinfo(T).max, n, dtype=T) if T is np.int64 else rng.random(n, dtype=np.float32)
    src = dab.B200Array.from_numpy(rt, a)
    out = dab.B200Array.empty(rt, (n,), T)
    tmp = dab.B200Array.empty(rt, (n,), T)
    _lib.call("dab_sort", rt.ctx, dab.dab_dtype(T), C.c_void_p(src.ptr), C.c_void_p(out.ptr), C.c_void_p(tmp.ptr), n)
    rt.sync()
    for b in (src, out, tmp):
        b.free()
dab.d_closeall()
rt.shutdown()
