"""One pass through the kernels added after the r1k capture (resident kernel, FK / Houdayer cluster moves, autocorrelation and
equilibration pushes) for an ncu launch list.  Usage: ncu_widened.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import peapods_b200 as pb  # noqa: E402

# C1 (README quickstart) through the resident kernel, then with Swendsen-Wang updates every sweep (the quickstart's --cluster-interval 1)
m = pb.Ising((32, 32), "ferro", np.linspace(1.5, 3.0, 16), n_replicas=2, seed=42)
m.sample(512, "metropolis", pt_interval=1)
m.sample(64, "metropolis", pt_interval=1, cluster_update_interval=1)
# critical 256^2 ferromagnet: labels in global scratch
m = pb.Ising((256, 256), "ferro", np.asarray([2.2, 2.269, 2.35]), n_replicas=2, seed=42)
m.sample(8, "metropolis", cluster_update_interval=1, cluster_mode="wolff")
# 3-D +-J spin glass, int8 layout: Metropolis + PT + Houdayer (README spin-glass recipe, 8^3)
m = pb.Ising((8, 8, 8), "bimodal", np.linspace(0.8, 1.4, 24), n_replicas=4, seed=42)
m.sample(32, "metropolis", pt_interval=1, overlap_cluster_update_interval=1)
# the same recipe on the multispin path at the headline geometry (1024 samples), with taus and the equilibration diagnostic
D = 1024
J = bench.make_couplings(0, D, D)
sim = pb.IsingSimulation(list(bench.SHAPE), J, bench.temperatures(), 4, None, bench.dynamics_seed(), layout="msc")
sim.sample(8, "metropolis", pt_interval=1, overlap_cluster_update_interval=1, autocorrelation_max_lag=2, equilibration_diagnostic=True,
           per_sample=False)
print("done")
