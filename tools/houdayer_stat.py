"""z-scores of <e>, <e^2>, <q^2> of one 4x4 +-J instance against exact enumeration, Metropolis + PT + Houdayer every sweep.
Usage: houdayer_stat.py [n_seeds] [n_sweeps]   (PP_LIB=path selects another build of the library)"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import peapods_b200 as pb  # noqa: E402
from peapods_b200 import _lib  # noqa: E402

if os.environ.get("PP_LIB"):
    _lib.LIB_PATH = Path(os.environ["PP_LIB"]).resolve()
from test_gpu_statistics import exact_2d_pm_j, zscores  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
base = int(sys.argv[3]) if len(sys.argv) > 3 else 700
cases = (("int8", "sw", 1),) if len(sys.argv) > 4 else (("int8", "sw", 1), ("int8", "wolff", 1), ("int8", "sw", None))
rng = np.random.default_rng(5)
J = (2 * rng.integers(0, 2, size=(4, 4, 2)) - 1).astype(np.float32)
temps = np.asarray([0.9, 1.6, 2.6], np.float32)
exact = exact_2d_pm_j(J.astype(np.float64), temps.astype(np.float64))
for layout, oc_mode, interval in cases:
    runs = []
    for seed in range(n_seeds):
        sim = pb.IsingSimulation([4, 4], J, temps, 2, None, base + seed, layout=layout)
        kw = dict(overlap_cluster_update_interval=interval, overlap_cluster_mode=oc_mode) if interval else {}
        r = sim.sample(n_sweeps, "metropolis", pt_interval=1, warmup_ratio=0.1, **kw)
        runs.append(np.stack([r["energies"], r["energies2"], r["overlap2"]], axis=1))
    runs = np.array(runs)
    z = zscores(runs, exact)
    print(layout, oc_mode, interval, "\nz=\n", np.round(z, 2), "\nmean=\n", runs.mean(axis=0), "\nexact=\n", exact, flush=True)
