#!/usr/bin/env python
"""Single-GPU rates of K11 (dab_sort: LSD radix sort of one chunk) and of sort(d::DVector) end to end.  Keys/s with CUDA events;
GB/s = elem * (1 + 2 * passes) * n / time, the onesweep kernel's algorithmic traffic (histogram read + per pass one read and one write)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import darray_b200 as dab  # noqa: E402
from darray_b200 import _lib  # noqa: E402

rt = dab.init(workers_per_rank=int(os.environ.get("WORKERS", "1")), use_dist=False)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = rt.event(), rt.event()
    rt.sync()
    rt.record(e0)
    for _ in range(reps):
        fn()
    rt.record(e1)
    return rt.elapsed_ms(e0, e1) / reps


rng = np.random.default_rng(1)
for name, T, n, gen, passes in (
        ("Int64 full range", np.int64, 1 << 27, lambda n: rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, n, dtype=np.int64), 8),
        ("Int64 in 0:10^6", np.int64, 1 << 27, lambda n: rng.integers(0, 10 ** 6, n, dtype=np.int64), 3),
        ("Float64 rand()", np.float64, 1 << 27, lambda n: rng.random(n), 7),
        ("Float32 rand()", np.float32, 1 << 28, lambda n: rng.random(n, dtype=np.float32), 4),
        ("Int32 full range", np.int32, 1 << 28, lambda n: rng.integers(np.iinfo(np.int32).min, np.iinfo(np.int32).max, n, dtype=np.int32), 4)):
    a = gen(n)
    src = dab.B200Array.from_numpy(rt, a)
    out = dab.B200Array.empty(rt, (n,), T)
    tmp = dab.B200Array.empty(rt, (n,), T)
    fn = lambda: _lib.call("dab_sort", rt.ctx, dab.dab_dtype(T), C.c_void_p(src.ptr), C.c_void_p(out.ptr), C.c_void_p(tmp.ptr), n)  # noqa: E731
    ms = timed(fn)
    es = np.dtype(T).itemsize
    got = out.to_numpy()
    ok = bool(np.all(got[:-1] <= got[1:])) and got[0] == a.min() and got[-1] == a.max()
    print(f"dab_sort {name:18s} n=2^{int(np.log2(n))} {ms:8.3f} ms {n / ms / 1e6:7.2f} Gkeys/s  ~{es * (1 + 2 * passes) * n / ms / 1e6:7.0f} GB/s "
          f"(<= {passes} passes)  sorted={ok}", flush=True)
    t0 = time.perf_counter()
    np.sort(a[: 1 << 24])
    t1 = time.perf_counter()
    print(f"         numpy sort of 2^24 on one host core: {(1 << 24) / (t1 - t0) / 1e9:.3f} Gkeys/s", flush=True)
    for b in (src, out, tmp):
        b.free()
    d = dab.distribute(a[: 1 << 26])
    ms = timed(lambda: dab.sort(d).close(), reps=3)
    print(f"sort(d::DVector) {name:18s} n=2^26, {len(d.chunks)} chunk(s): {ms:8.3f} ms {(1 << 26) / ms / 1e6:7.2f} Gkeys/s end to end", flush=True)
    d.close()
dab.d_closeall()
rt.shutdown()
