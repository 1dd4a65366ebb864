"""Where the time of one SlabIsingSimulation(...) + sample() goes (wall clock per phase). Usage: slab_e2e_prof.py [L] [sweeps]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
temps = np.asarray([4.511], np.float32)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = pb.IsingSimulation([L, L, L], "ferro", temps, 1, None, 7, layout="slab")
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    s.sample(n, "metropolis", warmup_ratio=0.25)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    del s
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"iter {it}: create {1e3 * (t1 - t0):.2f} ms, sample {1e3 * (t2 - t1):.2f} ms, destroy {1e3 * (t3 - t2):.2f} ms", flush=True)
