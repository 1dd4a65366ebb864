"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package; nothing under ``peapods_b200/`` does.  See ``pp_oracle.h`` for the
reference citations of every entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libpp_oracle.so"

RNG_XOSHIRO, RNG_PHILOX, RNG_PHILOX_MSC, RNG_PHILOX_PACKED, RNG_PHILOX_SYSQ = 0, 1, 2, 3, 4
SWEEP_METROPOLIS, SWEEP_GIBBS = 0, 1
PT_SINGLE_RANDOM_EDGE, PT_FULL_LADDER = 0, 1
TAG_INIT, TAG_SWEEP, TAG_PT, TAG_SWEEP_MSC = 0x00010000, 0x00020000, 0x00030000, 0x00040000
TAG_SWEEP_PACKED = 0x000A0000
TAG_SWEEP_SYSQ = 0x000B0000


def build(force: bool = False) -> Path:
    """Compile ``pp_oracle.c`` with the committed Makefile (gcc only)."""
    src_mtime = max((_HERE / f).stat().st_mtime for f in ("pp_oracle.c", "pp_oracle.h", "Makefile"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


class _Config(C.Structure):
    _fields_ = [
        ("n_sweeps", C.c_int64),
        ("warmup_sweeps", C.c_int64),
        ("sweep_mode", C.c_int32),
        ("pt_interval", C.c_int64),
        ("pt_schedule", C.c_int32),
        ("n_threads", C.c_int32),
        ("force_log_form", C.c_int32),
        ("autocorr_max_lag", C.c_int64),
        ("cluster_interval", C.c_int64),
        ("cluster_wolff", C.c_int32),
        ("overlap_cluster_interval", C.c_int64),
        ("overlap_cluster_wolff", C.c_int32),
        ("equil_diag", C.c_int32),
    ]


_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)


class _Results(C.Structure):
    _fields_ = (
        [(n, _PD) for n in ("mags", "mags2", "mags4", "energies", "energies2")]
        + [(n, _PD) for n in ("overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4")]
        + [("hist", _PU64), ("ql_at_q_sum", _PD), ("ql2_at_q_sum", _PD)]
        + [("ps_hist", _PU64), ("ps_ql_at_q_sum", _PD), ("ps_ql2_at_q_sum", _PD)]
        + [("edge_attempts", _PU64), ("edge_acceptances", _PU64), ("round_trips", _PU64)]
        + [("ps_means", _PD)]
        + [("mags2_tau", _PD), ("overlap2_tau", _PD), ("ps_taus", _PD)]
        + [("equil_energy_avg", _PD), ("equil_link_overlap_avg", _PD), ("ps_equil", _PD)]
    )


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(_LIB_PATH))
    vp, i64, i32, u64, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32
    f32 = C.c_float
    sig = {
        "orc_splitmix64": (u64, [u64]),
        "orc_child_seed": (u64, [u64, u64, u64]),
        "orc_realization_seed": (u64, [u64, u64]),
        "orc_xoshiro_seed_from_u64": (None, [vp, u64]),
        "orc_xoshiro_next_u64": (u64, [vp]),
        "orc_philox4x32_10": (None, [vp, vp, vp]),
        "orc_philox": (None, [vp, vp, vp]),
        "orc_philox4x32_r": (None, [vp, vp, C.c_int, vp]),
        "orc_draw24": (u32, [u64, u32, u32, u32, u32]),
        "orc_draw24_packed": (u32, [u64, u32, u32, u32, u32]),
        "orc_lattice_new": (vp, [i32, vp, i32, vp]),
        "orc_lattice_free": (None, [vp]),
        "orc_lattice_n_spins": (i64, [vp]),
        "orc_lattice_n_neighbors": (i32, [vp]),
        "orc_lattice_stride": (i64, [vp, i32]),
        "orc_neighbor_fwd": (u32, [vp, i64, i32]),
        "orc_neighbor_bwd": (u32, [vp, i64, i32]),
        "orc_colouring_is_valid": (i32, [vp, vp]),
        "orc_metropolis_lookup": (i32, [vp, i64, vp, i32, i32, vp]),
        "orc_metropolis_accepted_count": (u32, [f32, i32]),
        "orc_metropolis_legacy_accepts": (i32, [f32, i32, u32]),
        "orc_gibbs_accepted_count": (u32, [f32, i32]),
        "orc_gibbs_legacy_accepts": (i32, [f32, i32, u32]),
        "orc_energies_mags": (None, [vp, vp, vp, i64, vp, vp]),
        "orc_overlap_dots": (None, [vp, vp, vp, vp, vp]),
        "orc_sweep_xoshiro": (None, [vp, vp, vp, vp, vp, i64, vp, i32, i32]),
        "orc_sweep_philox": (None, [vp, vp, vp, vp, vp, i64, vp, u64, u32, i32, i32, i32]),
        "orc_full_ladder_edges": (i32, [i32, i32, vp]),
        "orc_pt_replay": (None, [i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "orc_sim_new": (vp, [i32, vp, i32, vp, vp, i64, vp, i32, i32, u64, i32, vp]),
        "orc_sim_free": (None, [vp]),
        "orc_sim_reset": (None, [vp, i32, u64]),
        "orc_sim_set_sample_offset": (None, [vp, i64]),
        "orc_sim_sample": (i32, [vp, C.POINTER(_Config), C.POINTER(_Results)]),
        "orc_sim_spins": (vp, [vp, i64]),
        "orc_sim_system_ids": (vp, [vp, i64]),
        "orc_sim_energies": (vp, [vp, i64]),
        "orc_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr, key):
    return philox4x32(ctr, key, 10)


def philox4x32(ctr, key, rounds=None):
    """Philox4x32-R; rounds=None: the RNG-SPEC generator (ORC_PHILOX_ROUNDS = 7)."""
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    if rounds is None:
        lib().orc_philox(_p(c), _p(k), _p(out))
    else:
        lib().orc_philox4x32_r(_p(c), _p(k), int(rounds), _p(out))
    return out


class Lattice:
    """geometry/lattice.rs:9-109"""

    def __init__(self, shape, offsets=None):
        self.shape = np.asarray(shape, dtype=np.int64)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        n_off = 0 if self.offsets is None else len(self.offsets)
        self.h = lib().orc_lattice_new(len(self.shape), _p(self.shape), n_off, _p(self.offsets))
        if not self.h:
            raise ValueError(lib().orc_last_error().decode())
        self.n_spins = lib().orc_lattice_n_spins(self.h)
        self.n_neighbors = lib().orc_lattice_n_neighbors(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lattice_free(self.h)
            self.h = None

    def strides(self):
        return [lib().orc_lattice_stride(self.h, d) for d in range(len(self.shape))]

    def fwd(self, i, d):
        return lib().orc_neighbor_fwd(self.h, i, d)

    def bwd(self, i, d):
        return lib().orc_neighbor_bwd(self.h, i, d)

    def fwd_table(self):
        return np.array([[self.fwd(i, d) for d in range(self.n_neighbors)] for i in range(self.n_spins)], dtype=np.uint32)

    def bwd_table(self):
        return np.array([[self.bwd(i, d) for d in range(self.n_neighbors)] for i in range(self.n_spins)], dtype=np.uint32)

    def colouring_is_valid(self, colour):
        c = np.ascontiguousarray(colour, dtype=np.uint16)
        return bool(lib().orc_colouring_is_valid(self.h, _p(c)))

    def energies_mags(self, spins, couplings):
        spins = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1, self.n_spins)
        J = np.ascontiguousarray(couplings, dtype=np.float32).reshape(-1)
        e = np.zeros(len(spins), dtype=np.float32)
        m = np.zeros(len(spins), dtype=np.int64)
        lib().orc_energies_mags(self.h, _p(spins), _p(J), len(spins), _p(e), _p(m))
        return e, m

    def overlap_dots(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.int8)
        b = np.ascontiguousarray(b, dtype=np.int8)
        ds, dl = C.c_int64(0), C.c_int64(0)
        lib().orc_overlap_dots(self.h, _p(a), _p(b), C.addressof(ds), C.addressof(dl))
        return ds.value, dl.value

    def sweep_xoshiro(self, spins, couplings, temperatures, system_ids, rng_states, sweep_mode, use_lookup):
        """In place on ``spins`` (int8 [S,N]) and ``rng_states`` (uint64 [S,4])."""
        J = np.ascontiguousarray(couplings, dtype=np.float32).reshape(-1)
        T = np.ascontiguousarray(temperatures, dtype=np.float32)
        sid = np.ascontiguousarray(system_ids, dtype=np.int64)
        assert spins.dtype == np.int8 and spins.flags.c_contiguous
        assert rng_states.dtype == np.uint64 and rng_states.flags.c_contiguous
        lib().orc_sweep_xoshiro(self.h, _p(spins), _p(J), _p(T), _p(sid), len(sid), _p(rng_states), sweep_mode, int(use_lookup))

    def sweep_philox(self, spins, couplings, temperatures, system_ids, colour, key, sweep_index, sweep_mode,
                     use_lookup=True, stream_is_slot=False):
        J = np.ascontiguousarray(couplings, dtype=np.float32).reshape(-1)
        T = np.ascontiguousarray(temperatures, dtype=np.float32)
        sid = np.ascontiguousarray(system_ids, dtype=np.int64)
        col = np.ascontiguousarray(colour, dtype=np.uint16)
        assert spins.dtype == np.int8 and spins.flags.c_contiguous
        lib().orc_sweep_philox(self.h, _p(spins), _p(J), _p(T), _p(sid), len(sid), _p(col), int(key), int(sweep_index),
                               sweep_mode, int(use_lookup), int(stream_is_slot))


def metropolis_lookup(couplings, temps, n_neighbors):
    """mcmc/sweep.rs:108-145; returns None when the reference returns None."""
    J = np.ascontiguousarray(couplings, dtype=np.float32).reshape(-1)
    T = np.ascontiguousarray(temps, dtype=np.float32)
    table = np.zeros((len(T), 4 * n_neighbors + 1), dtype=np.uint32)
    rc = lib().orc_metropolis_lookup(_p(J), len(J), _p(T), len(T), n_neighbors, _p(table))
    return None if rc != 0 else table


def full_ladder_edges(n_temps, first_parity):
    out = np.zeros(max(n_temps, 1), dtype=np.int32)
    n = lib().orc_full_ladder_edges(n_temps, first_parity, _p(out))
    return out[:n].tolist()


def pt_replay(n_replicas, temps, attempts):
    """attempts: list of (edge, accepted, left_system, right_system); realization.rs:73-120."""
    T = np.ascontiguousarray(temps, dtype=np.float32)
    n = len(attempts)
    e = np.array([a[0] for a in attempts], dtype=np.int32)
    acc = np.array([int(a[1]) for a in attempts], dtype=np.int32)
    le = np.array([a[2] for a in attempts], dtype=np.int64)
    ri = np.array([a[3] for a in attempts], dtype=np.int64)
    ea = np.zeros(max(len(T) - 1, 1), dtype=np.uint64)
    eacc = np.zeros(max(len(T) - 1, 1), dtype=np.uint64)
    rt = np.zeros(n_replicas * len(T), dtype=np.uint64)
    lib().orc_pt_replay(n_replicas, len(T), _p(T), n, _p(e), _p(acc), _p(le), _p(ri), _p(ea), _p(eacc), _p(rt))
    return ea[: len(T) - 1], eacc[: len(T) - 1], rt


class Sim:
    """simulation/mod.rs:865-939 driven like src/lib.rs:106-174 / 176-333 / 620-633."""

    def __init__(self, shape, couplings, temperatures, n_replicas=1, offsets=None, seed=42, rng_mode=RNG_XOSHIRO,
                 colour=None, sample_offset=0):
        self.shape = np.asarray(shape, dtype=np.int64)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        n_off = 0 if self.offsets is None else len(self.offsets)
        self.z = len(self.shape) if self.offsets is None else n_off
        self.N = int(np.prod(self.shape))
        J = np.ascontiguousarray(couplings, dtype=np.float32)
        single = tuple(int(s) for s in self.shape) + (self.z,)
        if J.shape == single:
            self.D = 1
        elif J.shape[1:] == single:
            self.D = J.shape[0]
        else:
            raise ValueError(f"couplings shape {J.shape} does not match lattice {single}")
        self.temps = np.ascontiguousarray(temperatures, dtype=np.float32)
        self.T, self.R = len(self.temps), int(n_replicas)
        self.colour = None if colour is None else np.ascontiguousarray(colour, dtype=np.uint16).reshape(-1)
        self.h = lib().orc_sim_new(len(self.shape), _p(self.shape), n_off, _p(self.offsets), _p(J), self.D, _p(self.temps),
                                   self.T, self.R, int(seed), rng_mode, _p(self.colour))
        if not self.h:
            raise ValueError(lib().orc_last_error().decode())
        if sample_offset:
            lib().orc_sim_set_sample_offset(self.h, int(sample_offset))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sim_free(self.h)
            self.h = None

    def reset(self, seed=None):
        lib().orc_sim_reset(self.h, int(seed is not None), 0 if seed is None else int(seed))

    def spins(self, realization=0):
        S = self.T * self.R
        ptr = lib().orc_sim_spins(self.h, realization)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int8)), shape=(S * self.N,)).copy()

    def system_ids(self, realization=0):
        ptr = lib().orc_sim_system_ids(self.h, realization)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(self.T * self.R,)).copy()

    def energies(self, realization=0):
        ptr = lib().orc_sim_energies(self.h, realization)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(self.T * self.R,)).copy()

    def sample(self, n_sweeps, sweep_mode="metropolis", pt_interval=None, pt_schedule="single_random_edge",
               warmup_ratio=0.25, n_threads=1, force_log_form=False, per_sample=True, autocorrelation_max_lag=None,
               equilibration_diagnostic=False, cluster_update_interval=None, cluster_mode="sw",
               overlap_cluster_update_interval=None, overlap_cluster_mode="wolff"):
        # src/lib.rs:219-220 (Rust f64::round = half away from zero)
        warm = int(np.floor(n_sweeps * warmup_ratio + 0.5))
        cfg = _Config(n_sweeps, warm, {"metropolis": 0, "gibbs": 1}[sweep_mode], 0 if pt_interval is None else pt_interval,
                      {"single_random_edge": 0, "full_ladder": 1}[pt_schedule], n_threads, int(force_log_form),
                      0 if autocorrelation_max_lag is None else int(autocorrelation_max_lag),
                      0 if cluster_update_interval is None else int(cluster_update_interval), {"sw": 0, "wolff": 1}[cluster_mode],
                      0 if overlap_cluster_update_interval is None else int(overlap_cluster_update_interval),
                      {"sw": 0, "wolff": 1}[overlap_cluster_mode],
                      int(bool(equilibration_diagnostic)))
        T, R, D, N = self.T, self.R, self.D, self.N
        out = {k: np.zeros(T, dtype=np.float64) for k in ("mags", "mags2", "mags4", "energies", "energies2")}
        res = _Results()
        keep = [out]
        for k in ("mags", "mags2", "mags4", "energies", "energies2"):
            setattr(res, k, out[k].ctypes.data_as(_PD))
        if R >= 2:
            for k in ("overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
                out[k] = np.zeros(T, dtype=np.float64)
                setattr(res, k, out[k].ctypes.data_as(_PD))
            out["overlap_histogram"] = np.zeros((T, N + 1), dtype=np.uint64)
            out["ql_at_q_sum"] = np.zeros((T, N + 1), dtype=np.float64)
            out["ql2_at_q_sum"] = np.zeros((T, N + 1), dtype=np.float64)
            res.hist = out["overlap_histogram"].ctypes.data_as(_PU64)
            res.ql_at_q_sum = out["ql_at_q_sum"].ctypes.data_as(_PD)
            res.ql2_at_q_sum = out["ql2_at_q_sum"].ctypes.data_as(_PD)
            if D > 1 and per_sample:
                out["per_sample_overlap_histogram"] = np.zeros((D, T, N + 1), dtype=np.uint64)
                out["per_sample_ql_at_q_sum"] = np.zeros((D, T, N + 1), dtype=np.float64)
                out["per_sample_ql2_at_q_sum"] = np.zeros((D, T, N + 1), dtype=np.float64)
                res.ps_hist = out["per_sample_overlap_histogram"].ctypes.data_as(_PU64)
                res.ps_ql_at_q_sum = out["per_sample_ql_at_q_sum"].ctypes.data_as(_PD)
                res.ps_ql2_at_q_sum = out["per_sample_ql2_at_q_sum"].ctypes.data_as(_PD)
        if pt_interval is not None:
            pt = {
                "edge_attempts": np.zeros((D, max(T - 1, 0)), dtype=np.uint64),
                "edge_acceptances": np.zeros((D, max(T - 1, 0)), dtype=np.uint64),
                "round_trips": np.zeros((D, R, T), dtype=np.uint64),
            }
            res.edge_attempts = pt["edge_attempts"].ctypes.data_as(_PU64)
            res.edge_acceptances = pt["edge_acceptances"].ctypes.data_as(_PU64)
            res.round_trips = pt["round_trips"].ctypes.data_as(_PU64)
            out["per_disorder"] = {"parallel_tempering": pt}
        self.last_per_sample_means = np.zeros((D, 11, T), dtype=np.float64)
        res.ps_means = self.last_per_sample_means.ctypes.data_as(_PD)
        if autocorrelation_max_lag is not None:  # src/lib.rs:545-556
            out["mags2_tau"] = np.zeros(T, dtype=np.float64)
            res.mags2_tau = out["mags2_tau"].ctypes.data_as(_PD)
            if R >= 2:
                out["overlap2_tau"] = np.zeros(T, dtype=np.float64)
                res.overlap2_tau = out["overlap2_tau"].ctypes.data_as(_PD)
            self.last_per_sample_taus = np.zeros((D, 2, T), dtype=np.float64)
            res.ps_taus = self.last_per_sample_taus.ctypes.data_as(_PD)
        if equilibration_diagnostic:  # src/lib.rs:559-574
            ck = equil_checkpoints(n_sweeps)
            out["equil_sweeps"] = np.asarray(ck, dtype=np.uint64)
            out["equil_energy_avg"] = np.zeros((len(ck), T), dtype=np.float64)
            out["equil_link_overlap_avg"] = np.zeros((len(ck), T), dtype=np.float64)
            res.equil_energy_avg = out["equil_energy_avg"].ctypes.data_as(_PD)
            res.equil_link_overlap_avg = out["equil_link_overlap_avg"].ctypes.data_as(_PD)
            self.last_per_sample_equil = np.zeros((D, len(ck), 2, T), dtype=np.float64)
            res.ps_equil = self.last_per_sample_equil.ctypes.data_as(_PD)
        rc = lib().orc_sim_sample(self.h, C.byref(cfg), C.byref(res))
        del keep
        if rc != 0:
            raise ValueError(lib().orc_last_error().decode())
        return out


def autocorr_gamma(values, max_lag):
    """AutocorrAccum (ring backend) over rows values[n_samples][n_temps] -> gamma[n_temps][max_lag + 1]
    (statistics/autocorrelation.rs:24-124, 166-199)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    if v.ndim == 1:
        v = v[:, None]
    n, T = v.shape if v.size else (0, v.shape[1] if v.ndim == 2 else 1)
    gamma = np.zeros((T, max_lag + 1), dtype=np.float64)
    f = lib().orc_autocorr_gamma
    f.restype = None
    f.argtypes = [_PD, C.c_int64, C.c_int, C.c_int, _PD]
    f(v.ctypes.data_as(_PD), n, T, max_lag, gamma.ctypes.data_as(_PD))
    return gamma


def sokal_tau(gamma):
    """statistics/autocorrelation.rs:201-210"""
    g = np.ascontiguousarray(gamma, dtype=np.float64)
    f = lib().orc_sokal_tau
    f.restype = C.c_double
    f.argtypes = [_PD, C.c_int]
    return float(f(g.ctypes.data_as(_PD), len(g)))


def equil_checkpoints(n_sweeps):
    """statistics/equilibration.rs:18-29"""
    f = lib().orc_equil_checkpoints
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.POINTER(C.c_int64)]
    buf = (C.c_int64 * 80)()
    n = f(int(n_sweeps), buf)
    return [int(buf[i]) for i in range(n)]
