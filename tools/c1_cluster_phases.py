"""Phase clocks of prows_cluster_resident_kernel at the quickstart (CTA 0, thread 0), from a library built with -DPP_PROWS_TIMING."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb  # noqa: E402
from peapods_b200 import _lib  # noqa: E402

if os.environ.get("PP_LIB"):
    _lib.LIB_PATH = Path(os.environ["PP_LIB"]).resolve()
m = pb.Ising((32, 32), "ferro", np.linspace(1.5, 3.0, 16), n_replicas=2, seed=42)
lib = _lib.load()
lib.pp_debug_prows_clocks.restype = C.c_int32
lib.pp_debug_prows_clocks.argtypes = [C.c_void_p, C.c_int32]
m.sample(256, "metropolis", pt_interval=1, warmup_ratio=0.0)
assert lib.pp_debug_prows_clocks(None, 1) == 1, "library was not built with -DPP_PROWS_TIMING"
n = 2048
m.sample(n, "metropolis", pt_interval=1, warmup_ratio=0.0)
buf = np.zeros(8, np.uint64)
assert lib.pp_debug_prows_clocks(buf.ctypes.data, 0) == 1
names = ["colour passes", "counts + all-gather", "cluster barrier A", "pair dots", "cluster barrier B", "fold", "exchange", "loop head"]
print(f"{n} recorded sweeps with an exchange each, loop {1e3 * m._sim.last_sweep_loop_ms / n:.2f} us per sweep; CTA 0 thread 0, us per sweep:")
for k in (7, 0, 1, 2, 3, 4, 5, 6):
    print(f"  {names[k]:22s} {buf[k] / n / 1965.0:7.3f}")
