#!/usr/bin/env python
"""Regenerates tests/golden/golden.json.

The reference is Julia and cannot be executed in this environment, so these are NOT outputs of the reference: they freeze
(1) the known-answer values the reference's own tests/docs state literally (copied with their file:line), and (2) the bit
patterns of the synthetic-input generator shared by the CUDA library and the oracle, so that neither side can drift silently.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import darray_oracle as orc  # noqa: E402

g = {
    "reference_literals": {
        "defaultdist_50_4": {"value": [1, 14, 27, 39, 51], "source": "test/darray.jl:66"},
        "sum_fill_1p1_100x100_local": {"value": "11000.000000000013", "source": "docs/src/index.md:222-225"},
        "sum_fill_1p1_100x100_distributed_8procs": {"value": "11000.000000000127", "source": "docs/src/index.md:227-230"},
        "prod_fill_2_10": {"value": 1024, "source": "test/darray.jl:512-518"},
        "reduce_fill_myid_10x10_two_procs": {"value": "50*MYID + 50*OTHERIDS", "source": "test/darray.jl:239-243"},
    },
    "rand_u01_f32_bits": {
        f"seed{seed}_start{start}": [int(v) for v in orc.rand_u01(seed, start, 16).view(np.uint32)]
        for seed, start in [(1234, 0), (1234, (1 << 32) + 5), (0, 0), (99, 12345)]
    },
    "rand_u01_ksum": {"seed1234_start0_n65536": orc.rand_u01_ksum(1234, 0, 65536)},
    "layouts": {
        f"{'x'.join(map(str, dims))}_np{npids}": {"grid": orc.defaultdist_grid(dims, npids),
                                                  "cuts": [orc.defaultdist_cuts(d, c) for d, c in zip(dims, orc.defaultdist_grid(dims, npids))]}
        for dims, npids in [((1024, 1024), 2), ((65536, 65536), 8), ((1 << 33,), 8), ((73, 73), 2), ((20, 20, 20), 8), ((50,), 4), ((100, 100), 6)]
    },
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
    json.dump(g, f, indent=1, sort_keys=True)
print("wrote golden.json")
