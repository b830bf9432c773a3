#!/usr/bin/env python
"""CPU-baseline table of BASELINE.md section 3: the oracle port (what the reference executes per worker, compute only) timed on
this box's host cores for P = 1, 2, 4, 8 and all cores.  GB/s = algorithmic bytes / mean pass time.  CPU only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import core as ocore  # noqa: E402

ocore.build()
cores = ocore.num_procs()
rows = [("map!(x->a*x+b, d, d)", 0, 8), ("sum(d)", 1, 4), ("maximum(d)", 2, 4), ("map! then sum (bench step)", 3, 12), ("sum(A, dims=1), 4096-row chunk", 4, 4),
        ("slab memcpy (upper bound for the TCP halo path)", 5, 4)]
out = {"host_cores": cores, "n_per_worker": 1 << 24, "rows": {}}
print(f"host cores online: {cores}; 2^24 Float32 (64 MiB) per worker; mean of 5 passes after 2 warm-ups")
print(f"{'operation':52s}" + "".join(f"{'P=' + str(p):>12s}" for p in (1, 2, 4, 8, cores)))
for name, op, bpe in rows:
    vals = []
    for P in (1, 2, 4, 8, cores):
        best, mean, _ = ocore.workers_run(op, P, 1 << 24, 1234, 1.5, 0.25, 2, 5)
        vals.append(bpe * (1 << 24) * P / mean / 1e9)
    out["rows"][name] = vals
    print(f"{name:52s}" + "".join(f"{v:12.1f}" for v in vals) + "   GB/s")
print(json.dumps(out))
