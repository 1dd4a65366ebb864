"""Where an end-to-end step (construct from host couplings -> sample -> result dict) spends its time."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench, peapods_b200 as pb
D = 4096
Jt = torch.empty((D,) + bench.SHAPE + (3,), dtype=torch.float32, pin_memory=True); J = Jt.numpy()
bench.make_couplings(0, D, D, out=J)
temps = bench.temperatures(); seed = bench.dynamics_seed()
kw = dict(pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25, per_sample=False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for it in range(6):
    t0 = time.perf_counter()
    s = pb.IsingSimulation(list(bench.SHAPE), J, temps, 4, None, seed, layout="msc")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = s.sample(n, "metropolis", **kw); t2 = time.perf_counter()
    loop = s.last_sweep_loop_ms
    del s; torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms, sample {1e3*(t2-t1):.1f} ms (sweep loop {loop:.1f}, overhead {1e3*(t2-t1)-loop:.1f}), destroy {1e3*(t3-t2):.1f}")
