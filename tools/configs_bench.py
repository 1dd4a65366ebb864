"""Throughput of the other BASELINE configs (C1, C3, C4) through the public API on one GPU.
They are parity-test cases, not bench.py lines; the numbers go into DESIGN.md.  Usage: configs_bench.py [c1 c3 c4]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os  # noqa: E402

import peapods_b200 as pb  # noqa: E402
from peapods_b200 import _lib  # noqa: E402

if os.environ.get("PP_LIB"):  # A/B runs against another build of the library
    _lib.LIB_PATH = Path(os.environ["PP_LIB"]).resolve()

TRI = [[1, 0], [0, 1], [1, -1]]


def run(name, model, n_sweeps, mode, reps=2, **kw):
    model.sample(max(4, n_sweeps // 4), mode, **kw)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        model.sample(n_sweeps, mode, **kw)
        wall = time.perf_counter() - t0
        dev = model._sim.last_sweep_loop_ms
        best = dev if best is None else min(best, dev)
    attempts = float(model.n_spins) * model.n_temps * model.n_replicas * model.n_disorder * n_sweeps
    print(f"{name}: layout={model._sim.layout} {n_sweeps} sweeps, sweep loop {best:.1f} ms (wall {1e3 * wall:.1f} ms), "
          f"{attempts / best / 1e6:.1f} attempts/ns, launches {model._sim.last_kernel_launches}", flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c3", "c4"]
    if "c1" in which:  # README quickstart: 2-D ferromagnet 32x32, Metropolis + PT, 16 temps, 2 replicas, 5000 sweeps
        m = pb.Ising((32, 32), "ferro", np.linspace(1.5, 3.0, 16), n_replicas=2, seed=42)
        run("C1 2-D Ising 32x32 T=16 R=2", m, 5000, "metropolis", pt_interval=1)
    if "c3" in which:  # triangular 256x256 via custom offsets, Gibbs, 64 temps around T_c = 4/ln 3, 2 replicas
        tc = 4.0 / np.log(3.0)
        m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), n_replicas=2, neighbor_offsets=TRI, seed=42)
        run("C3 triangular 256x256 T=64 R=2 Gibbs", m, 200, "gibbs")
    if "c4" in which:  # 3-D EA Gaussian 32^3, 48 temps, 4 replicas, 512 samples
        D = int(sys.argv[sys.argv.index("--c4-samples") + 1]) if "--c4-samples" in sys.argv else 512
        m = pb.Ising((32, 32, 32), "gaussian", np.linspace(0.8, 1.8, 48), n_replicas=4, n_disorder=D, seed=42)
        run(f"C4 3-D EA Gaussian 32^3 T=48 R=4 D={D}", m, 20, "metropolis", pt_interval=1, per_sample=False)


if __name__ == "__main__":
    main()
