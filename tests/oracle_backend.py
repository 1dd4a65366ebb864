"""A CPU stand-in with the ``peapods._core.IsingSimulation`` interface (src/lib.rs:106-174, 176-333, 620-633), backed by the oracle
in its reference-faithful mode.  TEST INFRASTRUCTURE: it lets the host-side layers (the reference's own ``run_sweep`` / CLI through
``peapods_b200.dropin``, and this repo's mirrors of them) run end to end in a container without a GPU."""
import numpy as np

import oracle


class OracleIsingSimulation:
    def __init__(self, lattice_shape, couplings, temperatures, n_replicas=None, neighbor_offsets=None, seed=None, **_engine_kwargs):
        shape = tuple(int(s) for s in lattice_shape)
        z = len(neighbor_offsets) if neighbor_offsets else len(shape)
        if isinstance(couplings, str):
            couplings = np.ones(shape + (z,), np.float32)
        self.layout = "oracle"
        self._sim = oracle.Sim(shape, np.asarray(couplings, np.float32), np.asarray(temperatures, np.float32),
                               n_replicas=1 if n_replicas is None else int(n_replicas), offsets=neighbor_offsets,
                               seed=42 if seed is None else int(seed), rng_mode=oracle.RNG_XOSHIRO)

    def sample(self, n_sweeps, sweep_mode, cluster_update_interval=None, cluster_mode=None, cluster_action=None, pt_interval=None,
               pt_schedule=None, overlap_cluster_update_interval=None, overlap_cluster_build_mode=None, overlap_cluster_mode=None,
               overlap_cluster_action=None, warmup_ratio=None, collect_cluster_stats=None, autocorrelation_max_lag=None,
               autocorrelation_backend=None, sequential=None, equilibration_diagnostic=None, snapshot_interval=None, **_engine_kwargs):
        if cluster_update_interval is not None or overlap_cluster_update_interval is not None or collect_cluster_stats or snapshot_interval:
            raise ValueError("the CPU stand-in runs single-spin-flip sweeps + parallel tempering only")
        res = self._sim.sample(int(n_sweeps), sweep_mode, pt_interval=pt_interval,
                               pt_schedule="single_random_edge" if pt_schedule is None else pt_schedule,
                               warmup_ratio=0.25 if warmup_ratio is None else float(warmup_ratio),
                               autocorrelation_max_lag=autocorrelation_max_lag,
                               equilibration_diagnostic=bool(equilibration_diagnostic))
        if "overlap_histogram" in res:  # src/lib.rs:369-383: a list of per-temperature arrays
            res["overlap_histogram"] = [h for h in res["overlap_histogram"]]
        return res

    def get_spins(self):
        return self._sim.spins(0)

    def reset(self, seed=None):
        self._sim.reset(seed)
