import sys, numpy as np
sys.path.insert(0, '/root/repo')
import peapods_b200 as pb
TRI = [[1, 0], [0, 1], [1, -1]]
tc = 4.0 / np.log(3.0)
m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), n_replicas=2, neighbor_offsets=TRI, seed=42)
m.sample(50, "gibbs")
for kw in (dict(), dict(warmup_ratio=1.0)):
    m.sample(200, "gibbs", profile=True, **kw)
    s = m._sim
    print(kw, f"loop {s.last_sweep_loop_ms:.2f} ms, sweep kernels {s.last_sweep_kernel_ms:.2f} ms over {s.last_sweep_kernel_launches} marks, launches {s.last_kernel_launches}")
    m.sample(200, "gibbs", **kw)
    print(kw, f"unprofiled loop {m._sim.last_sweep_loop_ms:.2f} ms")
