"""TEST INFRASTRUCTURE: run a script of the repo (bench.py, tools/*.py) with the host-memory emulation of the C ABI installed
(tests/hostmem_abi.py), i.e. exercise the script's HOST logic on a CPU-only machine:

    python tests/run_on_hostmem.py bench.py --log2n 16 --steps 2 --warmup 3 --no-cpu --no-extras

Numbers printed by such a run are meaningless (the "kernels" are NumPy); what it checks is that the script still drives the public API
correctly end to end -- e.g. that bench.py's in-run parity block comes out true."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import darray_b200  # noqa: E402,F401
import hostmem_abi  # noqa: E402

lib_mod = sys.modules["darray_b200._lib"]
bc_mod = sys.modules["darray_b200._broadcast"]
fake = hostmem_abi.HostMemABI()
lib_mod._lib = fake
_real_codegen = bc_mod.codegen


def _recording_codegen(e):
    src = _real_codegen(e)
    fake.exprs[src.encode()] = e
    return src


bc_mod.codegen = _recording_codegen
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
runpy.run_path(os.path.join(ROOT, script), run_name="__main__")
