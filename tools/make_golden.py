"""Golden vectors from the REFERENCE's own Python host layer (python/peapods/spin_models.py), generated in the build
container where /root/reference is mounted.  The Rust extension `peapods._core` cannot be built here (no cargo), so it is
replaced by a recording stub: what is pinned is everything the reference computes ABOVE the extension boundary —
seed derivation, coupling generation, the argument lists handed to `IsingSimulation(...)`, `.sample(...)`, `.reset(...)`,
and the post-processed observables — i.e. the inputs and outputs of the drop-in boundary.

    python tools/make_golden.py            # writes tests/golden/host_layer.npz + host_layer.json

tests/test_golden_host_layer.py replays the same calls through peapods_b200.spin_models.Ising with the same stub and
compares bit for bit.  Nothing here runs on the GPU box (it has no /root/reference)."""
import hashlib
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference/python")

CASES = [  # name, constructor kwargs
    ("sq8_bimodal", dict(lattice_shape=(8, 8), couplings="bimodal", temperatures=np.linspace(1.0, 3.0, 5), n_replicas=2, seed=42)),
    ("cube4_bimodal_d3", dict(lattice_shape=(4, 4, 4), couplings="bimodal", temperatures=np.linspace(0.8, 1.4, 4), n_replicas=4,
                              n_disorder=3, seed=7)),
    ("tri6_gauss_d2", dict(lattice_shape=(6, 6), couplings="gaussian", temperatures=np.geomspace(0.5, 4.0, 6), n_replicas=2,
                           n_disorder=2, geometry="triangular", seed=123)),
    ("fcc4_ferro", dict(lattice_shape=(4, 4, 4), couplings="ferro", temperatures=np.linspace(8.0, 11.0, 3), geometry="fcc", seed=5)),
    ("bcc4_bimodal", dict(lattice_shape=(4, 4, 4), couplings="bimodal", temperatures=np.linspace(2.0, 8.0, 3), n_replicas=2,
                          geometry="bcc", seed=0)),
    ("c2_l16_bimodal_d4", dict(lattice_shape=(16, 16, 16), couplings="bimodal", temperatures=np.linspace(0.8, 1.4, 32), n_replicas=4,
                               n_disorder=4, seed=42)),  # BASELINE config 2 geometry (first 4 of its realizations)
    ("custom_offsets", dict(lattice_shape=(5, 7), couplings="gaussian", temperatures=np.linspace(0.8, 2.5, 2), n_replicas=2,
                            neighbor_offsets=[[1, 0], [0, 1], [1, 1]], seed=2**40 + 3)),
]
SAMPLE_CALLS = [  # kwargs of Ising.sample per case index (cycled)
    dict(n_sweeps=100),
    dict(n_sweeps=57, sweep_mode="gibbs", pt_interval=3, pt_schedule="full_ladder", warmup_ratio=0.1),
    dict(n_sweeps=10, pt_interval=1),
]


def synthetic_result(n_temps, n_spins, with_overlap, seed):
    """A result dict of the extension's shape (src/lib.rs:337-412) filled from a seeded generator."""
    rng = np.random.default_rng(seed)
    r = {k: rng.random(n_temps) + 0.1 for k in ("mags", "mags2", "mags4", "energies", "energies2")}
    if with_overlap:
        for k in ("overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
            r[k] = rng.random(n_temps) + 0.1
        r["overlap_histogram"] = [rng.integers(0, 50, size=n_spins + 1).astype(np.uint64) for _ in range(n_temps)]
    return r


class Recorder:
    def __init__(self):
        self.calls = []

    def make_core(self):
        rec = self

        class IsingSimulation:
            def __init__(self, lattice_shape, couplings, temperatures, n_replicas=None, neighbor_offsets=None, seed=None, **kw):
                self.n_temps, self.n_spins, self.n_replicas = len(temperatures), int(np.prod(lattice_shape)), n_replicas or 1
                rec.calls.append(("init", dict(lattice_shape=list(lattice_shape), couplings=couplings, temperatures=temperatures,
                                               n_replicas=n_replicas, neighbor_offsets=neighbor_offsets, seed=seed)))

            def sample(self, n_sweeps, sweep_mode, **kw):
                rec.calls.append(("sample", dict(n_sweeps=n_sweeps, sweep_mode=sweep_mode, **kw)))
                return synthetic_result(self.n_temps, self.n_spins, self.n_replicas >= 2, seed=n_sweeps)

            def reset(self, seed=None):
                rec.calls.append(("reset", dict(seed=seed)))

        return IsingSimulation


def run_cases(ising_cls, recorder):
    """Drive `ising_cls` (the reference's or ours) through CASES; returns {case: {array name: array}} and the call log."""
    out = {}
    for i, (name, kw) in enumerate(CASES):
        recorder.calls.clear()
        model = ising_cls(**kw)
        init = recorder.calls[0][1]
        coup = np.asarray(init["couplings"], dtype=np.float32)
        arrays = {
            "couplings_sha256": np.frombuffer(hashlib.sha256(np.ascontiguousarray(coup).tobytes()).digest(), dtype=np.uint8),
            "couplings_shape": np.asarray(coup.shape, dtype=np.int64),
            "couplings_head": coup.reshape(-1)[:64].copy(),
            "temperatures": np.asarray(init["temperatures"]),
            "dynamics_seed": np.asarray([init["seed"]], dtype=np.uint64),
        }
        skw = SAMPLE_CALLS[i % len(SAMPLE_CALLS)]
        model.sample(**skw)
        for attr in ("binder_cumulant", "heat_capacity", "sg_binder", "link_overlap_binder"):
            if hasattr(model, attr):
                arrays[attr] = np.asarray(getattr(model, attr), dtype=np.float64)
        model.reset()
        model.reset(seed=1000 + i)
        log = []
        for kind, args in recorder.calls:
            a = {k: v for k, v in args.items() if k not in ("couplings", "temperatures")}
            if kind == "init":
                a["seed"] = int(a["seed"])
            if kind == "reset" and a["seed"] is not None:
                a["seed"] = int(a["seed"])
            log.append([kind, a])
        out[name] = (arrays, log)
    return out


def main():
    if not REFERENCE.exists():
        raise SystemExit("the reference tree is not mounted: golden vectors can only be regenerated in the build container")
    recorder = Recorder()
    pkg = types.ModuleType("peapods")
    pkg.__path__ = [str(REFERENCE / "peapods")]
    core = types.ModuleType("peapods._core")
    core.IsingSimulation = recorder.make_core()
    sys.modules["peapods"] = pkg
    sys.modules["peapods._core"] = core
    import importlib

    ref = importlib.import_module("peapods.spin_models")  # the reference's own file, unmodified
    res = run_cases(ref.Ising, recorder)
    seeds = [0, 1, 42, 12345, 2**32 + 1, 2**63 - 1]
    arrays = {"seed_list": np.asarray(seeds, dtype=np.uint64),
              "dynamics_seeds": np.asarray([ref._dynamics_seed(s) for s in seeds], dtype=np.uint64)}
    logs = {}
    for name, (arrs, log) in res.items():
        for k, v in arrs.items():
            arrays[f"{name}/{k}"] = v
        logs[name] = log
    gold = ROOT / "tests" / "golden"
    gold.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(gold / "host_layer.npz", **arrays)
    (gold / "host_layer.json").write_text(json.dumps(logs, indent=1, sort_keys=True))
    print("wrote", gold / "host_layer.npz", (gold / "host_layer.npz").stat().st_size, "bytes,", len(arrays), "arrays")


if __name__ == "__main__":
    main()
