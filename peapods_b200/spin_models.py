"""``Ising`` — the user-facing model class of the reference
(/root/reference/python/peapods/spin_models.py:26-362) over the B200 engine.

Same constructor and ``sample`` keywords, same attributes after sampling, same error texts.
Coupling generation follows spin_models.py:13-19 and :104-126 exactly (numpy ``SeedSequence``:
child 0 seeds the couplings, child 1 the dynamics; one grandchild per disorder realization), so a
given ``seed`` produces the couplings the reference produces."""
from __future__ import annotations

import numpy as np

from ._core import IsingSimulation

GEOMETRIES = {  # spin_models.py:5-10
    "triangular": [[1, 0], [0, 1], [1, -1]],
    "tri": [[1, 0], [0, 1], [1, -1]],
    "fcc": [[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1]],
    "bcc": [[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1]],
}


def seed_material(seed):
    """(coupling SeedSequence, dynamics u64) for a user seed — spin_models.py:13-19."""
    if seed is not None and (not isinstance(seed, (int, np.integer)) or seed < 0):
        raise ValueError("seed must be a non-negative integer or None")
    coupling_seq, dynamics_seq = np.random.SeedSequence(seed).spawn(2)
    return coupling_seq, int(dynamics_seq.generate_state(1, dtype=np.uint64)[0])


def make_couplings(kind, lattice_shape, n_neighbors, n_disorder, coupling_seq):
    """spin_models.py:107-126: one child generator per realization, realization 0 is stable when
    ``n_disorder`` grows (tests/test_sampling_interfaces.py:45-48)."""
    single = tuple(lattice_shape) + (n_neighbors,)
    per_realization = []
    for child in coupling_seq.spawn(n_disorder):
        rng = np.random.default_rng(child)
        if kind == "ferro":
            block = np.ones(single, dtype=np.float32)
        elif kind == "bimodal":
            block = (2 * rng.integers(0, 2, size=single) - 1).astype(np.float32)
        elif kind == "gaussian":
            block = rng.standard_normal(single).astype(np.float32)
        else:
            raise ValueError("Invalid mode for couplings.")
        per_realization.append(block)
    return per_realization[0] if n_disorder == 1 else np.stack(per_realization)


class Ising:
    """Ising model on a periodic Bravais lattice, sampled on the GPU sweep engine."""

    def __init__(self, lattice_shape, couplings="ferro", temperatures=np.geomspace(0.1, 10, 32), n_replicas=1,
                 n_disorder=1, neighbor_offsets=None, geometry=None, seed=None, *, layout="auto", device=0, **engine_kwargs):
        if geometry is not None:
            if neighbor_offsets is not None:
                raise ValueError("Cannot specify both geometry and neighbor_offsets")
            if geometry not in GEOMETRIES:
                raise ValueError(f"Unknown geometry '{geometry}', choose from: {list(GEOMETRIES.keys())}")
            neighbor_offsets = GEOMETRIES[geometry]
        self.lattice_shape = tuple(lattice_shape)
        self.n_spins = int(np.prod(lattice_shape))
        self.n_dims = len(lattice_shape)
        self.n_neighbors = len(neighbor_offsets) if neighbor_offsets else self.n_dims
        self.temperatures = np.asarray(temperatures).copy().astype(np.float32)
        self.n_temps = len(temperatures)
        self.n_replicas = n_replicas
        self.n_disorder = n_disorder
        self.seed = seed
        coupling_seq, self._constructor_dynamics_seed = seed_material(seed)
        if isinstance(couplings, np.ndarray):
            coup = couplings.astype(np.float32)
        elif couplings == "ferro" and n_disorder == 1 and self.n_spins * self.n_neighbors >= (1 << 26):
            coup = "ferro"  # a large ferromagnet's all-ones array (12 GiB at 1024^3) is never materialised
        else:
            coup = make_couplings(couplings, self.lattice_shape, self.n_neighbors, n_disorder, coupling_seq)
        self.couplings = coup
        self._neighbor_offsets = neighbor_offsets
        self._layout_request, self._device, self._engine_kwargs = layout, device, dict(engine_kwargs)
        self._sim = IsingSimulation(list(lattice_shape), coup, self.temperatures, n_replicas, neighbor_offsets,
                                    self._constructor_dynamics_seed, layout=layout, device=device, **engine_kwargs)

    def _fall_back_to_int8(self):
        """The multispin layout (chosen by ``layout="auto"`` for >= 32 +-J realizations) has no cluster labels per lane, so
        Swendsen-Wang / Wolff updates and SW-mode Houdayer moves run on the int8 layout only.  A model on the automatic layout
        that asks for them is moved there: same configurations and system ids, a fresh handle (its sweep / exchange counters
        start again, and its draws are the int8 layout's per-realization streams)."""
        old = self._sim
        new = IsingSimulation(list(self.lattice_shape), self.couplings, self.temperatures, self.n_replicas, self._neighbor_offsets,
                              self._constructor_dynamics_seed, layout="int8", device=self._device, **self._engine_kwargs)
        for d in range(self.n_disorder):
            new.set_system_ids(old.get_system_ids(d), d)
            new.set_spins(old.get_spins(d), d)
        self._sim = new

    def reset(self, seed=None):
        """Replay the constructor's dynamics, or a one-off seeded reset (spin_models.py:138-144)."""
        self._sim.reset(None if seed is None else seed_material(seed)[1])

    def sample(self, n_sweeps, sweep_mode="metropolis", cluster_update_interval=None, cluster_mode="sw",
               cluster_action="update", pt_interval=None, pt_schedule="single_random_edge",
               overlap_cluster_update_interval=None, overlap_cluster_build_mode="houdayer",
               overlap_cluster_mode="wolff", overlap_cluster_action="update", warmup_ratio=0.25,
               collect_cluster_stats=False, autocorrelation_max_lag=None, autocorrelation_backend="ring",
               sequential=False, equilibration_diagnostic=False, snapshot_interval=None, **engine_kwargs):
        # argument checks in the reference's order and wording (spin_models.py:222-247)
        if cluster_action not in {"update", "observe"}:
            raise ValueError("cluster_action must be 'update' or 'observe'")
        if overlap_cluster_action not in {"update", "observe"}:
            raise ValueError("overlap_cluster_action must be 'update' or 'observe'")
        if pt_schedule not in {"single_random_edge", "full_ladder"}:
            raise ValueError("pt_schedule must be 'single_random_edge' or 'full_ladder'")
        if autocorrelation_backend not in {"ring", "fft"}:
            raise ValueError("autocorrelation_backend must be 'ring' or 'fft'")
        if autocorrelation_backend == "fft" and autocorrelation_max_lag is None:
            raise ValueError("autocorrelation_backend='fft' requires autocorrelation_max_lag")
        if cluster_action == "observe" and cluster_update_interval is None:
            raise ValueError("cluster_action='observe' requires cluster_update_interval")
        if overlap_cluster_action == "observe" and overlap_cluster_update_interval is None:
            raise ValueError("overlap_cluster_action='observe' requires overlap_cluster_update_interval")

        oci = overlap_cluster_update_interval
        if (self._layout_request == "auto" and getattr(self._sim, "layout", None) == "msc"
                and (cluster_update_interval is not None or (oci is not None and overlap_cluster_mode == "sw"))
                and cluster_action == "update" and overlap_cluster_action == "update" and not collect_cluster_stats
                and overlap_cluster_build_mode.strip() in ("houdayer", "houd2") and snapshot_interval is None):
            self._fall_back_to_int8()
        result = self._sim.sample(
            n_sweeps, sweep_mode,
            cluster_update_interval=cluster_update_interval,
            cluster_mode=cluster_mode if cluster_update_interval else None,
            cluster_action=cluster_action if cluster_update_interval else None,
            pt_interval=pt_interval, pt_schedule=pt_schedule,
            overlap_cluster_update_interval=oci,
            overlap_cluster_build_mode=overlap_cluster_build_mode if oci else None,
            overlap_cluster_mode=overlap_cluster_mode if oci else None,
            overlap_cluster_action=overlap_cluster_action if oci else None,
            warmup_ratio=warmup_ratio, collect_cluster_stats=collect_cluster_stats,
            autocorrelation_max_lag=autocorrelation_max_lag, autocorrelation_backend=autocorrelation_backend,
            sequential=sequential, equilibration_diagnostic=equilibration_diagnostic,
            snapshot_interval=snapshot_interval if oci else None, **engine_kwargs)

        # post-processing, spin_models.py:270-293
        self.mags, self.mags2, self.mags4 = result["mags"], result["mags2"], result["mags4"]
        self.energies_avg, self.energies2_avg = result["energies"], result["energies2"]
        with np.errstate(divide="ignore", invalid="ignore"):
            self.binder_cumulant = 1 - self.mags4 / (3 * self.mags2**2)
            self.heat_capacity = self.n_spins * (self.energies2_avg - self.energies_avg**2) / self.temperatures**2
            if "overlap2" in result:
                self.overlap, self.overlap2, self.overlap4 = result["overlap"], result["overlap2"], result["overlap4"]
                self.sg_binder = 1 - self.overlap4 / (3 * self.overlap2**2)
                self.link_overlap = result["link_overlap"]
                self.link_overlap2 = result["link_overlap2"]
                self.link_overlap4 = result["link_overlap4"]
                self.link_overlap_binder = 1 - self.link_overlap4 / (3 * self.link_overlap2**2)
        for key in ("overlap_histogram", "ql_at_q_sum", "ql2_at_q_sum", "per_sample_overlap_histogram",
                    "per_sample_ql_at_q_sum", "per_sample_ql2_at_q_sum"):
            if key in result:
                setattr(self, key, result[key])
        if "mags2_tau" in result:  # spin_models.py:311-314
            self.mags2_tau = result["mags2_tau"]
        if "overlap2_tau" in result:
            self.overlap2_tau = result["overlap2_tau"]
        if "equil_sweeps" in result:  # spin_models.py:316-319
            self._equil_sweeps = result["equil_sweeps"]
            self._equil_energy_avg = result["equil_energy_avg"]
            self._equil_link_overlap_avg = result["equil_link_overlap_avg"]
        self.per_disorder = result.get("per_disorder", {})
        return result

    def equilibration_delta(self, j_squared=1.0):
        """Delta(t) = e(t) - J^2 beta z (1 - q_l(t)) at the checkpoints (spin_models.py:322-341; sign convention
        e = +sum J s s / N).  Returns (sweeps [n_ckpt], delta [n_ckpt, n_temps])."""
        beta = 1.0 / self.temperatures
        delta = self._equil_energy_avg - j_squared * beta * self.n_neighbors * (1 - self._equil_link_overlap_avg)
        return self._equil_sweeps, delta

    def get_energies(self):
        """Mean energy per temperature of the last run (sign: e = +sum J s s / N, spin_models.py:343-344)."""
        return self.energies_avg
