"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py [family ...]

Families: msc3d (bulk-async staging on mbarriers, named barriers, fire-and-forget reductions), msc (table-driven multispin),
rows (row-table int8: ferro / +-J / fp32, in-sweep energies with last-block finalisation), resident (one CTA per realization),
int8 (site tables, odd extents), slab (bytes), slabp (one bit per spin), fk (label propagation), houdayer (int8 + multispin flood
fill), stats (autocorrelation ring, equilibration checkpoints, parallel tempering)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb  # noqa: E402

TRI = [[1, 0], [0, 1], [1, -1]]
rng = np.random.default_rng(1)


def pmj(D, shape, z):
    return (2 * rng.integers(0, 2, size=(D,) + tuple(shape) + (z,)) - 1).astype(np.float32)


def run(name, sim, n=6, mode="metropolis", **kw):
    kw.setdefault("pt_interval", 1)
    r = sim.sample(n, mode, warmup_ratio=0.25, **kw)
    assert np.all(np.isfinite(r["energies"]))
    sim.get_spins(0)
    print(f"sanitize[{name}]: layout={sim.layout} ok", flush=True)


CASES = {
    "msc3d": lambda: run("msc3d", pb.IsingSimulation([8, 8, 8], pmj(32, (8, 8, 8), 3), np.linspace(0.8, 1.4, 4).astype(np.float32), 4, None, 3, layout="msc")),
    "msc3d_fl": lambda: run("msc3d full ladder gibbs", pb.IsingSimulation([4, 4, 8], pmj(40, (4, 4, 8), 3), np.linspace(0.8, 1.4, 3).astype(np.float32), 2, None, 3, layout="msc"),
                            mode="gibbs", pt_schedule="full_ladder"),
    "msc": lambda: run("msc generic", pb.IsingSimulation([6, 6], pmj(32, (6, 6), 3), np.linspace(1.0, 2.0, 3).astype(np.float32), 2, TRI, 3, layout="msc")),
    "rows": lambda: [run("rows ferro tri gibbs", pb.IsingSimulation([16, 16], "ferro", np.linspace(3.2, 4.0, 4).astype(np.float32), 2, TRI, 3, layout="int8"), mode="gibbs"),
                     run("rows +-J", pb.IsingSimulation([4, 4, 8], pmj(3, (4, 4, 8), 3), np.linspace(0.8, 1.4, 3).astype(np.float32), 2, None, 3, layout="int8")),
                     run("rows f32", pb.IsingSimulation([4, 4, 8], rng.standard_normal((3, 4, 4, 8, 3)).astype(np.float32), np.linspace(0.8, 1.4, 3).astype(np.float32), 2, None, 3, layout="int8"))],
    "resident": lambda: run("resident", pb.IsingSimulation([16, 16], "ferro", np.linspace(1.5, 3.0, 4).astype(np.float32), 2, None, 3, layout="int8"), n=40),
    "int8": lambda: run("int8 tables (odd extents)", pb.IsingSimulation([5, 7], pmj(2, (5, 7), 2), np.linspace(1.0, 2.0, 3).astype(np.float32), 2, None, 3, layout="int8")),
    "slab": lambda: run("slab bytes", pb.IsingSimulation([8, 4, 16], "ferro", np.asarray([4.0, 4.5], np.float32), 1, None, 3, layout="slab", slab_ranks=2, slab_rank=-1)),
    "slabp": lambda: [run("slab bits", pb.IsingSimulation([8, 4, 64], "ferro", np.asarray([4.0, 4.5], np.float32), 1, None, 3, layout="slab", slab_ranks=2, slab_rank=-1)),
                      run("slab bits gibbs", pb.IsingSimulation([4, 4, 128], "ferro", np.asarray([4.0, 4.5, 1e9], np.float32), 1, None, 3, layout="slab"), mode="gibbs")],
    "fk": lambda: run("fk", pb.IsingSimulation([8, 8], "ferro", np.linspace(2.0, 2.6, 3).astype(np.float32), 2, None, 3, layout="int8"), cluster_update_interval=1, cluster_mode="sw"),
    "houdayer": lambda: [run("houdayer int8", pb.IsingSimulation([4, 4, 4], pmj(2, (4, 4, 4), 3), np.linspace(0.8, 1.4, 3).astype(np.float32), 4, None, 3, layout="int8"),
                             overlap_cluster_update_interval=1, overlap_cluster_mode="sw"),
                         run("houdayer msc", pb.IsingSimulation([4, 4, 8], pmj(32, (4, 4, 8), 3), np.linspace(0.8, 1.4, 3).astype(np.float32), 2, None, 3, layout="msc"),
                             overlap_cluster_update_interval=1)],
    "stats": lambda: run("autocorr + equil", pb.IsingSimulation([4, 4, 8], pmj(2, (4, 4, 8), 3), np.linspace(0.8, 1.4, 3).astype(np.float32), 2, None, 3, layout="int8"),
                         n=24, autocorrelation_max_lag=3, equilibration_diagnostic=True, pt_schedule="full_ladder"),
}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
    print("SANITIZE_CASES DONE")
