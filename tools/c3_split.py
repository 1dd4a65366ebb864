"""C3 (triangular 256^2 ferromagnet, Gibbs, 64 temperatures, 2 replicas): device time per sweep of warm-up-only batches (pure
sweeps, up to 64 per launch) against recorded sweeps (sweep + energies launch, overlap + fold launch)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb  # noqa: E402

TRI = [[1, 0], [0, 1], [1, -1]]
tc = 4.0 / np.log(3.0)
m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), n_replicas=2, neighbor_offsets=TRI, seed=42)
for label, kw in (("pure sweeps (warmup_ratio=1)", dict(warmup_ratio=1.0)), ("recorded sweeps (warmup_ratio=0)", dict(warmup_ratio=0.0)),
                  ("recorded + PT every sweep", dict(warmup_ratio=0.0, pt_interval=1))):
    m.sample(64, "gibbs", **kw)
    best = None
    for _ in range(3):
        m.sample(256, "gibbs", **kw)
        dev = m._sim.last_sweep_loop_ms
        best = dev if best is None else min(best, dev)
    att = 256.0 * 256 * 128 * 256
    print(f"{label}: {1e3 * best / 256:.2f} us per sweep, {att / best / 1e6:.1f} attempts/ns, launches {m._sim.last_kernel_launches}", flush=True)
