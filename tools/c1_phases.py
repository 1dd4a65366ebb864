import sys, numpy as np
sys.path.insert(0, '/root/repo')
import peapods_b200 as pb
m = pb.Ising((32, 32), "ferro", np.linspace(1.5, 3.0, 16), n_replicas=2, seed=42)
for name, kw in [("pure sweeps", dict(warmup_ratio=1.0)), ("+PT", dict(warmup_ratio=1.0, pt_interval=1)),
                 ("recorded, no PT", dict(warmup_ratio=0.0)), ("recorded + PT", dict(warmup_ratio=0.0, pt_interval=1))]:
    m.sample(500, "metropolis", **kw)
    m.sample(2000, "metropolis", **kw)
    print(f"{name}: {m._sim.last_sweep_loop_ms / 2000 * 1e3:.2f} us/sweep")
