"""Small C3 / C4 runs for ncu captures of the int8 row-table kernels. Usage: ncu_small.py c3|c4 [D]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb
TRI = [[1, 0], [0, 1], [1, -1]]
if sys.argv[1] == "c3":
    tc = 4.0 / np.log(3.0)
    m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), n_replicas=2, neighbor_offsets=TRI, seed=42)
    m.sample(6, "gibbs")
else:
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    m = pb.Ising((32, 32, 32), "gaussian", np.linspace(0.8, 1.8, 48), n_replicas=4, n_disorder=D, seed=42)
    m.sample(4, "metropolis", pt_interval=1, per_sample=False)
print("done", m._sim.last_sweep_loop_ms)
