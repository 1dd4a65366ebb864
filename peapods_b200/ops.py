"""The reference's operator-level functions with their HOST-SLICE signatures, over the C ABI's ``pp_slice_*`` entry points
(include/peapods_b200.h): each call ships its slices to the device, runs one kernel and brings the result back.

    metropolis_sweep / gibbs_sweep                     spin-sim/src/mcmc/sweep.rs:220-229, 262-270
    compute_energies_and_magnetizations                spin-sim/src/spins/energy.rs:59-65
    overlap_dots  (OverlapAccum::collect)              spin-sim/src/statistics/overlap.rs:251-281
    parallel_tempering / parallel_tempering_full_ladder  spin-sim/src/mcmc/tempering.rs:20-27, 45-53

Unit-level parity only: a simulation keeps a handle (``IsingSimulation``) and never moves spins per sweep.
Where the reference takes ``rngs`` the draws here are RNG-SPEC's, a pure function of (seed, sweep_index | pt_event)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ModelDesc


def _desc(lattice_shape, couplings, temperatures, n_replicas, neighbor_offsets, seed, layout, device):
    shape = np.asarray([int(s) for s in lattice_shape], dtype=np.int64)
    offsets = None if neighbor_offsets is None else np.ascontiguousarray(neighbor_offsets, dtype=np.int64)
    temps = np.ascontiguousarray(temperatures, dtype=np.float32).reshape(-1)
    d = ModelDesc()
    keep = [shape, offsets, temps]
    if isinstance(couplings, str):
        if couplings != "ferro":
            raise ValueError("couplings must be an array or 'ferro'")
        d.coupling_kind, d.couplings = 1, None
    else:
        z = len(shape) if offsets is None else len(offsets)
        coup = np.ascontiguousarray(couplings, dtype=np.float32)
        if coup.shape != tuple(int(s) for s in shape) + (z,):  # src/lib.rs:146-149
            raise ValueError(f"couplings shape {list(coup.shape)} does not match lattice {list(shape) + [z]}")
        d.coupling_kind, d.couplings = 0, coup.ctypes.data
        keep.append(coup)
    d.n_dims, d.shape = len(shape), shape.ctypes.data
    d.n_offsets = 0 if offsets is None else len(offsets)
    d.offsets = None if offsets is None else offsets.ctypes.data
    d.n_disorder, d.sample_offset = 1, 0
    d.temperatures, d.n_temps, d.n_replicas = temps.ctypes.data, len(temps), int(n_replicas)
    d.seed = int(seed)
    d.layout, d.device = _lib.LAYOUTS[layout], int(device)
    d.slab_ranks, d.slab_rank, d.nccl_unique_id = 1, 0, None
    d.system_ranks, d.system_rank = 1, 0
    return d, keep, int(np.prod(shape)), len(temps) * int(n_replicas)


def _sweep(mode, lattice_shape, spins, couplings, temperatures, system_ids, n_replicas, neighbor_offsets, seed, sweep_index, exact_log,
           layout, device):
    d, keep, N, S = _desc(lattice_shape, couplings, temperatures, n_replicas, neighbor_offsets, seed, layout, device)
    if not (isinstance(spins, np.ndarray) and spins.dtype == np.int8 and spins.flags.c_contiguous and spins.size == S * N):
        raise ValueError("spins must be a C-contiguous int8 array of n_systems * n_spins entries (updated in place)")
    ids = None if system_ids is None else np.ascontiguousarray(system_ids, dtype=np.int64).reshape(-1)
    _lib.check(_lib.load().pp_slice_sweep(C.byref(d), _lib.SWEEP_MODES[mode], int(sweep_index), int(bool(exact_log)), spins.ctypes.data,
                                          None if ids is None else ids.ctypes.data))
    del keep


def metropolis_sweep(lattice_shape, spins, couplings, temperatures, system_ids=None, *, n_replicas=1, neighbor_offsets=None, seed=42,
                     sweep_index=0, exact_log=True, layout="int8", device=0):
    """One checkerboard Metropolis sweep of every system, in place (mcmc/sweep.rs:220-229).  ``temperatures``: the T ladder values,
    systems are slot-major ``replica * T + slot``; ``system_ids[slot]`` picks the configuration a slot sweeps (parallel.rs:27-33)."""
    _sweep("metropolis", lattice_shape, spins, couplings, temperatures, system_ids, n_replicas, neighbor_offsets, seed, sweep_index,
           exact_log, layout, device)


def gibbs_sweep(lattice_shape, spins, couplings, temperatures, system_ids=None, *, n_replicas=1, neighbor_offsets=None, seed=42,
                sweep_index=0, exact_log=True, layout="int8", device=0):
    """One checkerboard heat-bath sweep of every system, in place (mcmc/sweep.rs:262-270)."""
    _sweep("gibbs", lattice_shape, spins, couplings, temperatures, system_ids, n_replicas, neighbor_offsets, seed, sweep_index, exact_log,
           layout, device)


def compute_energies_and_magnetizations(lattice_shape, spins, couplings, n_systems, *, neighbor_offsets=None, layout="int8", device=0):
    """(energies f32 [n_systems], magnetisation sums i64 [n_systems]) — spins/energy.rs:59-65, 99-108."""
    d, keep, N, S = _desc(lattice_shape, couplings, np.ones(int(n_systems), np.float32), 1, neighbor_offsets, 0, layout, device)
    s = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1)
    if s.size != S * N:
        raise ValueError("spins must hold n_systems * n_spins entries")
    e, m = np.zeros(S, np.float32), np.zeros(S, np.int64)
    _lib.check(_lib.load().pp_slice_energies_mags(C.byref(d), s.ctypes.data, e.ctypes.data, m.ctypes.data))
    del keep
    return e, m


def overlap_dots(lattice_shape, spins, temperatures, system_ids=None, *, n_replicas=2, neighbor_offsets=None, layout="int8", device=0):
    """(dot_spin, dot_link) i64 [n_replicas // 2, T] of the replica pairs (2p, 2p + 1) at every temperature slot
    (statistics/overlap.rs:259-281); the couplings play no part."""
    d, keep, N, S = _desc(lattice_shape, "ferro", temperatures, n_replicas, neighbor_offsets, 0, layout, device)
    s = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1)
    if s.size != S * N:
        raise ValueError("spins must hold n_systems * n_spins entries")
    ids = None if system_ids is None else np.ascontiguousarray(system_ids, dtype=np.int64).reshape(-1)
    T = S // int(n_replicas)
    ds, dl = np.zeros((int(n_replicas) // 2, T), np.int64), np.zeros((int(n_replicas) // 2, T), np.int64)
    _lib.check(_lib.load().pp_slice_overlap(C.byref(d), s.ctypes.data, None if ids is None else ids.ctypes.data, ds.ctypes.data, dl.ctypes.data))
    del keep
    return ds, dl


def _pt(schedule, lattice_shape, energies, temperatures, system_ids, n_replicas, seed, pt_event, first_parity, device):
    d, keep, N, S = _desc(lattice_shape, "ferro", temperatures, n_replicas, None, seed, "int8", device)
    e = np.ascontiguousarray(energies, dtype=np.float32).reshape(-1)
    ids = np.ascontiguousarray(system_ids, dtype=np.int64).reshape(-1).copy()
    if e.size != S or ids.size != S:
        raise ValueError("energies and system_ids must have n_replicas * n_temps entries")
    _lib.check(_lib.load().pp_slice_pt(C.byref(d), _lib.PT_SCHEDULES[schedule], int(pt_event), int(first_parity), e.ctypes.data, ids.ctypes.data))
    del keep
    return ids


def parallel_tempering(lattice_shape, energies, temperatures, system_ids, *, n_replicas=1, seed=42, pt_event=0, device=0):
    """One single-random-edge exchange attempt per replica ladder (mcmc/tempering.rs:20-42): returns the new ``system_ids``.
    ``energies`` are per SYSTEM (energy per spin, as compute_energies returns them); the lattice only supplies n_spins."""
    return _pt("single_random_edge", lattice_shape, energies, temperatures, system_ids, n_replicas, seed, pt_event, 0, device)


def parallel_tempering_full_ladder(lattice_shape, energies, temperatures, system_ids, *, n_replicas=1, seed=42, pt_event=0, first_parity=0,
                                   device=0):
    """Every edge of one parity, then of the other (mcmc/tempering.rs:45-70)."""
    return _pt("full_ladder", lattice_shape, energies, temperatures, system_ids, n_replicas, seed, pt_event, first_parity, device)
