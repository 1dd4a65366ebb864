"""Quick A/B timing of pp_sample on the C2 workload (device time of the sweep loop). Usage: kbench.py [D] [sweeps] [reps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import os
import bench, peapods_b200 as pb
from peapods_b200 import _lib
if os.environ.get("PP_LIB"):
    _lib.LIB_PATH = Path(os.environ["PP_LIB"]).resolve()

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
warm = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
pt = None if os.environ.get("KB_NOPT") else 1
J = bench.make_couplings(0, D, D)
sim = pb.IsingSimulation(list(bench.SHAPE), J, bench.temperatures(), 4, None, bench.dynamics_seed(), layout="msc")
kw = dict(pt_interval=pt, pt_schedule="single_random_edge", warmup_ratio=warm, per_sample=False)
if os.environ.get("KB_HOUDAYER"):  # Houdayer move every KB_HOUDAYER sweeps (the reference's spin-glass recipe)
    kw["overlap_cluster_update_interval"] = int(os.environ["KB_HOUDAYER"])
for _ in range(2):
    sim.sample(n, "metropolis", **kw)
ms = []
for _ in range(reps):
    sim.sample(n, "metropolis", **kw)
    ms.append(sim.last_sweep_loop_ms)
best = min(ms)
att = 4096.0 * 128 * D * n
print(f"D={D} sweeps={n} warm={warm} pt={pt}: best {best:.2f} ms  {best/n*1e3:.1f} us/sweep  {att/best/1e6:.0f} attempts/ns   all={['%.2f' % m for m in ms]}")
