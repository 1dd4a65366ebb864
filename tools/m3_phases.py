"""Phase clocks of msc3d_kernel CTAs (variant build with -DPP_M3_TIMING, loaded through PP_LIB): where a CTA's lifetime goes.
Usage: PP_LIB=build/libpp_m3timing.so PP_STREAMS=1 python tools/m3_phases.py [D] [sweeps] [warm]"""
import ctypes as C
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import bench
import peapods_b200 as pb
from peapods_b200 import _lib

_lib.LIB_PATH = Path(os.environ["PP_LIB"]).resolve()
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
warm = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
J = bench.make_couplings(0, D, D)
sim = pb.IsingSimulation(list(bench.SHAPE), J, bench.temperatures(), 4, None, bench.dynamics_seed(), layout="msc")
kw = dict(pt_interval=None if os.environ.get("KB_NOPT") else 1, pt_schedule="single_random_edge", warmup_ratio=warm, per_sample=False)
for _ in range(2):
    sim.sample(n, "metropolis", **kw)
print(f"loop {sim.last_sweep_loop_ms / n * 1e3:.1f} us per sweep")
lib = _lib.load()
lib.pp_debug_m3_clocks.restype = C.c_int32
lib.pp_debug_m3_clocks.argtypes = [C.c_void_p, C.c_int64]
ncta = min(8192, (D // 32) * 32)
buf = np.zeros(8192 * 32, np.uint64)
assert lib.pp_debug_m3_clocks(buf.ctypes.data, buf.size) == 0
c = buf.reshape(8192, 32)[:ncta].astype(np.int64)
mhz = 1.965e3  # SM clock (cycles per us)
names = ["init -> loads issued / gather merged", "-> data ready", "-> colour 0 done", "-> colour 1 done", "-> stage-out issued",
         "-> epilogue done", "-> smem reads of the store done"]
dt = np.diff(c[:, :8], axis=1) / mhz
life = (c[:, 7] - c[:, 0]) / mhz
print(f"{ncta} CTAs of the last launch; CTA lifetime (clock64) mean {life.mean():.2f} us, median {np.median(life):.2f}, p90 {np.percentile(life, 90):.2f}")
for k, nm in enumerate(names):
    print(f"  {nm:42s} mean {dt[:, k].mean():6.2f} us   median {np.median(dt[:, k]):6.2f}   p90 {np.percentile(dt[:, k], 90):6.2f}")
has = c[:, 14] > c[:, 0]
if has.any():
    h = c[has]
    print(f"  stage-in detail ({has.mean():.2f} of the CTAs gather): masks arrived +{((h[:, 12] - h[:, 0]) / mhz).mean():.2f} us, gather loads issued "
          f"+{((h[:, 13] - h[:, 12]) / mhz).mean():.2f}, bulk spins waited +{((h[:, 14] - h[:, 13]) / mhz).mean():.2f}, merged +{((h[:, 1] - h[:, 14]) / mhz).mean():.2f}")
print(f"  barrier init + sync done (all CTAs) +{((c[:, 15] - c[:, 0]) / mhz).mean():.2f} us; bulk copies issued, masks arrived +{((c[:, 12] - c[:, 0]) / mhz).mean():.2f} us")
def at(k):
    return ((c[:, k] - c[:, 0]) / mhz).mean()
print("  thread 0, us after kernel start: bulk copies issued %.2f | small loads issued %.2f | sync %.2f | masks arrived %.2f"
      % (at(18), at(16), at(15), at(12)))
print("  gather loads issued %.2f | bulk spins arrived %.2f | merged, ids parked %.2f | J + spins waited, barrier %.2f" % (at(13), at(14), at(1), at(2)))
print("  epilogue, us after stage-out issue: counters parked + barrier %.2f | energy merge %.2f | pair walk %.2f | parked + barrier %.2f | pair merges + barrier %.2f | replica tail %.2f | end %.2f"
      % tuple(((c[:, k] - c[:, 5]) / mhz).mean() for k in (22, 23, 24, 25, 26, 27, 6)))
gt0, gt1, sm = c[:, 10], c[:, 9], c[:, 11]
print(f"launch span (globaltimer) {(gt1.max() - gt0.min()) / 1e3:.1f} us; lifetimes by globaltimer mean {((gt1 - gt0) / 1e3).mean():.2f} us")
# per-SM occupancy: sum of CTA lifetimes on an SM / span, and the gaps between a CTA's end and the next start on the same SM
busy, gaps = [], []
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    busy.append(((gt1[idx] - gt0[idx]).sum()) / max(1, (gt1.max() - gt0.min())))
    order = idx[np.argsort(gt0[idx])]
    ends = np.sort(gt1[idx])
    st = np.sort(gt0[idx])
    # k-th start (after the first two resident CTAs) follows the (k-2)-th end
    if len(st) > 2:
        gaps.extend(((st[2:] - ends[:-2]) / 1e3).tolist())
print(f"SMs used {len(busy)}; mean resident CTAs per SM over the launch span {np.mean(busy):.2f}; end -> next start on the SM: mean {np.mean(gaps):.2f} us, median {np.median(gaps):.2f}")
