"""Phase clocks of prows_sweep_kernel in cluster mode at C3 (CTA 0, thread 0), from a library built with -DPP_PROWS_TIMING:
  nvcc ... -DPP_PROWS_TIMING -o peapods_b200/lib/variants/libpp_prows_timing.so peapods_b200/csrc/pp_engine.cu
  PP_LIB=peapods_b200/lib/variants/libpp_prows_timing.so python tools/c3_phases.py [R]"""
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
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TRI = [[1, 0], [0, 1], [1, -1]]
tc = 4.0 / np.log(3.0)
m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 128 // R), n_replicas=R, neighbor_offsets=TRI, seed=42)
lib = _lib.load()
lib.pp_debug_prows_clocks.restype = C.c_int32
lib.pp_debug_prows_clocks.argtypes = [C.c_void_p, C.c_int32]
m.sample(64, "gibbs", warmup_ratio=0.0)
assert lib.pp_debug_prows_clocks(None, 1) == 1, "library was not built with -DPP_PROWS_TIMING"
n = 256
m.sample(n, "gibbs", warmup_ratio=0.0)
buf = np.zeros(8, np.uint64)
assert lib.pp_debug_prows_clocks(buf.ctypes.data, 0) == 1
names = ["colour passes", "energy phase", "cluster barrier 1", "pair dots", "cluster barrier 2", "fold", "-", "loop head"]
print(f"R = {R}, {n} recorded sweeps, loop {1e3 * m._sim.last_sweep_loop_ms / n:.2f} us per sweep; CTA 0 thread 0, us per sweep at 1.965 GHz:")
for k in (7, 0, 1, 2, 3, 4, 5):
    print(f"  {names[k]:20s} {buf[k] / n / 1965.0:7.3f}")
