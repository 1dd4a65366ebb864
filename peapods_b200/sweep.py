"""``run_sweep`` and its ``.npz`` writer — the reference's parameter-sweep driver
(/root/reference/python/peapods/sweep.py:100-163 ``_save_data``, :351-512 ``run_sweep``) over the B200 engine.

Same keyword list, same seed derivation (one ``SeedSequence`` of the user seed, one child per (coupling kind, shape):
sweep.py:11-27), same labels, same ``sweep_<config>.npz`` keys, shapes and dtypes, so files written here load in whatever reads
the reference's files.  Plots are not part of this package (matplotlib is the reference's optional extra): ``save_plots=True``
raises.  The reference's own, unmodified ``run_sweep`` / CLI also run on this engine through ``peapods_b200.dropin``."""
from __future__ import annotations

import itertools
import sys
import time
from pathlib import Path

import numpy as np

COUPLING_SEED_TAGS = {"ferro": 0, "bimodal": 1, "gaussian": 2}  # sweep.py:10


def run_seed_words(seed):
    """sweep.py:13-19"""
    if seed is not None and (not isinstance(seed, (int, np.integer)) or seed < 0):
        raise ValueError("seed must be a non-negative integer or None")
    return [int(v) for v in np.random.SeedSequence(seed).generate_state(4, dtype=np.uint32)]


def run_child_seed(root_words, coupling, shape):
    """sweep.py:22-27: the model seed of one (coupling kind, lattice shape) run"""
    seq = np.random.SeedSequence(root_words, spawn_key=(COUPLING_SEED_TAGS[coupling], len(shape), *shape))
    return int(seq.generate_state(1, dtype=np.uint64)[0])


def flatten_per_disorder_arrays(per_disorder, prefix=""):
    """sweep.py:30-45"""
    flat = {}
    key_prefix = f"{prefix}_" if prefix else ""
    for kind, fields in per_disorder.get("cluster_observations", {}).items():
        for field, values in fields.items():
            flat[f"{key_prefix}per_disorder_cluster_observations_{kind}_{field}"] = values
    pt = per_disorder.get("parallel_tempering")
    if pt is not None:
        for field, values in pt.items():
            flat[f"{key_prefix}per_disorder_pt_{field}"] = values
    return flat


def cumulative_overlap_ratio(per_sample_hist):
    """I(q)/X(q) of Billoire et al. from per-sample overlap histograms ``[D, T, bins]`` — sweep.py:48-79.
    Returns (q grid, median / mean ratio ``[T, n_q]``, mean, median)."""
    n_disorder, n_temps, n_bins = per_sample_hist.shape
    centers = np.linspace(-1, 1, n_bins)
    center = n_bins // 2
    positive = centers[center:]
    n_q = len(positive)
    x = np.zeros((n_disorder, n_temps, n_q))
    for qi in range(n_q):
        x[:, :, qi] = per_sample_hist[:, :, center - qi:center + qi + 1].sum(axis=2)
    totals = per_sample_hist.sum(axis=2, keepdims=True)
    x /= np.where(totals == 0, 1, totals)
    mean, median = x.mean(axis=0), np.median(x, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(mean > 0, median / mean, 0.0)
    return positive, ratio, mean, median


def config_label(coupling, build_mode, oc_mode):
    """sweep.py:82-88"""
    parts = [coupling]
    if build_mode != "houdayer":
        parts.append(build_mode)
    if oc_mode != "wolff":
        parts.append(oc_mode)
    return "_".join(parts)


def size_label(shape):
    return "x".join(str(s) for s in shape)


def save_sweep_data(models, label, temperatures, output_dir):
    """``sweep_<label>.npz`` with the reference's keys (sweep.py:100-163); returns the path."""
    d = {"temperatures": temperatures}
    for prefix, model in models.items():
        d[f"{prefix}_lattice_shape"] = np.array(model.lattice_shape)
        d[f"{prefix}_binder_cumulant"] = model.binder_cumulant
        d[f"{prefix}_heat_capacity"] = model.heat_capacity
        d[f"{prefix}_energies"] = model.energies_avg
        if hasattr(model, "sg_binder"):
            d[f"{prefix}_sg_binder"] = model.sg_binder
        if hasattr(model, "overlap_histogram"):
            d[f"{prefix}_overlap_histogram"] = np.array([h for h in model.overlap_histogram])
        if hasattr(model, "per_sample_overlap_histogram"):
            d[f"{prefix}_per_sample_overlap_histogram"] = model.per_sample_overlap_histogram
            q_grid, ratio, _, _ = cumulative_overlap_ratio(model.per_sample_overlap_histogram)
            d[f"{prefix}_cumulative_overlap_q"] = q_grid
            d[f"{prefix}_cumulative_overlap_ratio"] = ratio
        if hasattr(model, "mags2_tau"):
            d[f"{prefix}_mags2_tau"] = model.mags2_tau
        if hasattr(model, "overlap2_tau"):
            d[f"{prefix}_overlap2_tau"] = model.overlap2_tau
        if hasattr(model, "_equil_sweeps"):
            d[f"{prefix}_equil_sweeps"] = model._equil_sweeps
            d[f"{prefix}_equil_energy_avg"] = model._equil_energy_avg
            d[f"{prefix}_equil_link_overlap_avg"] = model._equil_link_overlap_avg
        d.update(flatten_per_disorder_arrays(model.per_disorder, prefix=prefix))
    path = Path(output_dir) / f"sweep_{label}.npz"
    np.savez(path, **d)
    print(f"  Data saved to {path}")
    return path


def run_sweep(sizes, *, couplings=("ferro",), temperatures, n_replicas=1, n_disorder=1, neighbor_offsets=None, geometry=None,
              n_sweeps, sweep_mode="metropolis", cluster_update_interval=None, cluster_mode="sw", cluster_action="update",
              pt_interval=None, pt_schedule="single_random_edge", overlap_cluster_update_interval=None,
              overlap_cluster_build_modes=("houdayer",), overlap_cluster_modes=("wolff",), overlap_cluster_action="update",
              warmup_ratio=0.25, collect_cluster_stats=False, autocorrelation_max_lag=None, autocorrelation_backend="ring",
              autocorrelation_plot_temp=None, equilibration_diagnostic=False, save_plots=False, save_data=False, output_dir=".",
              sequential=False, snapshot_interval=None, seed=None, model_cls=None, **model_kwargs):
    """Cartesian sweep over (coupling kind, overlap build mode, overlap cluster mode) x sizes — sweep.py:351-512.
    Returns ``{config label: {size label: Ising}}``; with ``save_data`` one ``sweep_<config>.npz`` per config."""
    if save_plots:
        raise ValueError("save_plots needs the reference's matplotlib helpers; write the data (save_data=True) and plot from the .npz")
    if model_cls is None:
        from .spin_models import Ising as model_cls
    if save_data:
        Path(output_dir).mkdir(parents=True, exist_ok=True)
    valid = []
    for coupling, build_mode, oc_mode in itertools.product(couplings, overlap_cluster_build_modes, overlap_cluster_modes):
        if build_mode != "houdayer" and overlap_cluster_update_interval is None:  # sweep.py:95-101
            print(f"  skip: {config_label(coupling, build_mode, oc_mode)} — overlap_cluster_build_mode={build_mode} set but no "
                  f"--overlap-cluster-update-interval", file=sys.stderr)
            continue
        valid.append((coupling, build_mode, oc_mode))
    total, run_idx, results = len(valid) * len(sizes), 0, {}
    wall_start = time.perf_counter()
    seed_words = run_seed_words(seed)
    for coupling, build_mode, oc_mode in valid:
        label = config_label(coupling, build_mode, oc_mode)
        models = {}
        for shape in sizes:
            run_idx += 1
            slabel = size_label(shape)
            print(f"[{run_idx}/{total}] {slabel}, {label}")
            model = model_cls(shape, couplings=coupling, temperatures=temperatures, n_replicas=n_replicas, n_disorder=n_disorder,
                              neighbor_offsets=neighbor_offsets, geometry=geometry, seed=run_child_seed(seed_words, coupling, shape),
                              **model_kwargs)
            t0 = time.perf_counter()
            model.sample(n_sweeps, sweep_mode=sweep_mode, cluster_update_interval=cluster_update_interval, cluster_mode=cluster_mode,
                         cluster_action=cluster_action, pt_interval=pt_interval, pt_schedule=pt_schedule,
                         overlap_cluster_update_interval=overlap_cluster_update_interval, overlap_cluster_build_mode=build_mode,
                         overlap_cluster_mode=oc_mode, overlap_cluster_action=overlap_cluster_action, warmup_ratio=warmup_ratio,
                         collect_cluster_stats=collect_cluster_stats, autocorrelation_max_lag=autocorrelation_max_lag,
                         autocorrelation_backend=autocorrelation_backend, sequential=sequential,
                         equilibration_diagnostic=equilibration_diagnostic, snapshot_interval=snapshot_interval)
            print(f"  {time.perf_counter() - t0:.2f}s")
            models[slabel] = model
        results[label] = models
        if save_data:
            save_sweep_data(models, label, temperatures, output_dir)
    print(f"\nSweep complete: {total} runs in {time.perf_counter() - wall_start:.1f}s")
    return results
