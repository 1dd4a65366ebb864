import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import peapods_b200 as pb
TRI = [[1, 0], [0, 1], [1, -1]]
tc = 4.0 / np.log(3.0)
for R in (1, 2, 4):
    m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 128 // R), n_replicas=R, neighbor_offsets=TRI, seed=42)
    for label, kw in (("pure", dict(warmup_ratio=1.0)), ("recorded", dict(warmup_ratio=0.0))):
        m.sample(64, "gibbs", **kw)
        best = None
        for _ in range(3):
            m.sample(256, "gibbs", **kw)
            best = m._sim.last_sweep_loop_ms if best is None else min(best, m._sim.last_sweep_loop_ms)
        print(f"R={R} {label}: {1e3 * best / 256:.2f} us per sweep, launches {m._sim.last_kernel_launches}", flush=True)
