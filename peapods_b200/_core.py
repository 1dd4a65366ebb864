"""``IsingSimulation`` — the binding class the reference exposes as ``peapods._core.IsingSimulation``
(PyO3, /root/reference/src/lib.rs:12-20, 106-174, 176-333, 620-633), re-expressed over the C ABI of
``libpeapods_b200.so``.  Same constructor / ``sample`` / ``get_spins`` / ``reset`` signatures, same
result-dict keys, shapes and dtypes (src/lib.rs:337-490); options whose code paths are outside the
GPU sweep engine raise ``ValueError`` before any state is touched."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ModelDesc, Results, SampleCfg

_F64_KEYS = _lib.RESULT_F64


def _rust_round(x: float) -> int:
    """f64::round — half away from zero (src/lib.rs:220)."""
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


class IsingSimulation:
    def __init__(self, lattice_shape, couplings, temperatures, n_replicas=None, neighbor_offsets=None, seed=None,
                 *, layout="auto", device=0, sample_offset=0, slab_ranks=1, slab_rank=0, nccl_unique_id=None,
                 system_ranks=1, system_rank=0):
        """``layout="slab"`` (or "auto" for one large ferromagnet) keeps ONE 3-D lattice without neighbour tables and
        can cut it along dimension 0 into ``slab_ranks`` slabs, one per process / GPU: pass this process's
        ``slab_rank`` and the bytes of ``nccl_unique_id()`` from rank 0 (``slab_rank=-1`` keeps every slab on this
        device)."""
        lib = _lib.load()
        self._lib = lib
        self._h = C.c_void_p()
        self.lattice_shape = [int(s) for s in lattice_shape]
        shape = np.asarray(self.lattice_shape, dtype=np.int64)
        offsets = None if neighbor_offsets is None else np.ascontiguousarray(neighbor_offsets, dtype=np.int64)
        if offsets is not None and (offsets.ndim != 2 or offsets.shape[1] != len(shape)):
            raise ValueError(f"offset has length {offsets.shape[-1] if offsets.ndim else 0}, expected {len(shape)}")
        self.n_neighbors = len(shape) if offsets is None else len(offsets)
        self.n_spins = int(np.prod(shape))
        self.n_replicas = 1 if n_replicas is None else int(n_replicas)
        temps = np.ascontiguousarray(temperatures, dtype=np.float32).reshape(-1)
        self.n_temps = len(temps)
        expected_single = tuple(self.lattice_shape) + (self.n_neighbors,)

        desc = ModelDesc()
        if isinstance(couplings, str):
            if couplings != "ferro":
                raise ValueError("couplings must be an array or 'ferro'")
            desc.coupling_kind = 1
            desc.couplings = None
            self.n_realizations = 1
            coup = None
        else:
            coup = np.ascontiguousarray(couplings, dtype=np.float32)
            if coup.shape == expected_single:
                self.n_realizations = 1
            elif coup.ndim == len(expected_single) + 1 and coup.shape[1:] == expected_single:
                self.n_realizations = coup.shape[0]
            else:  # src/lib.rs:146-149
                raise ValueError(f"couplings shape {list(coup.shape)} does not match lattice {list(expected_single)}")
            desc.coupling_kind = 0
            desc.couplings = coup.ctypes.data
        desc.n_dims = len(shape)
        desc.shape = shape.ctypes.data
        desc.n_offsets = 0 if offsets is None else len(offsets)
        desc.offsets = None if offsets is None else offsets.ctypes.data
        desc.n_disorder = self.n_realizations
        desc.sample_offset = int(sample_offset)
        desc.temperatures = temps.ctypes.data
        desc.n_temps = self.n_temps
        desc.n_replicas = self.n_replicas
        desc.seed = 42 if seed is None else int(seed)  # src/lib.rs:155
        desc.layout = _lib.LAYOUTS[layout]
        desc.device = int(device)
        desc.slab_ranks = int(slab_ranks)
        desc.slab_rank = int(slab_rank)
        desc.system_ranks = int(system_ranks)
        desc.system_rank = int(system_rank)
        id_buf = None
        if nccl_unique_id is not None:
            id_buf = np.frombuffer(bytes(nccl_unique_id), dtype=np.uint8).copy()
            if id_buf.size != _lib.NCCL_ID_BYTES:
                raise ValueError(f"nccl_unique_id must have {_lib.NCCL_ID_BYTES} bytes")
            desc.nccl_unique_id = id_buf.ctypes.data
        _lib.check(lib.pp_create(C.byref(desc), C.byref(self._h)))
        self.n_local_spins = int(lib.pp_local_spin_count(self._h))
        self.layout = _lib.LAYOUT_NAMES[lib.pp_get_layout(self._h)]
        self.uses_msc3d = bool(lib.pp_uses_msc3d(self._h))
        self.slab_packed = bool(lib.pp_slab_packed(self._h))
        self.rows_packed = bool(lib.pp_rows_packed(self._h))
        self.sys_words = bool(lib.pp_sys_words(self._h))
        self.last_sweep_loop_ms = 0.0
        self.last_kernel_launches = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.pp_destroy(h)
            self._h = C.c_void_p()

    # ------------------------------------------------------------------ sample (src/lib.rs:176-618)
    def sample(self, n_sweeps, sweep_mode, cluster_update_interval=None, cluster_mode=None, cluster_action=None,
               pt_interval=None, pt_schedule=None, overlap_cluster_update_interval=None, overlap_cluster_build_mode=None,
               overlap_cluster_mode=None, overlap_cluster_action=None, warmup_ratio=None, collect_cluster_stats=None,
               autocorrelation_max_lag=None, autocorrelation_backend=None, sequential=None,
               equilibration_diagnostic=None, snapshot_interval=None, *, exact_log=False, per_sample=True,
               on_sweep=None, interrupt=None, profile=False):
        warmup = 0.25 if warmup_ratio is None else float(warmup_ratio)
        n_sweeps = int(n_sweeps)
        if sweep_mode not in _lib.SWEEP_MODES:  # config.rs:9-20
            raise ValueError(f"unknown sweep_mode '{sweep_mode}', expected 'metropolis' or 'gibbs'")
        schedule = "single_random_edge" if pt_schedule is None else pt_schedule  # src/lib.rs:225
        if schedule not in _lib.PT_SCHEDULES:  # config.rs:67-79
            raise ValueError(f"unknown pt_schedule '{schedule}', expected 'single_random_edge' or 'full_ladder'")
        backend = "ring" if autocorrelation_backend is None else autocorrelation_backend
        if backend not in ("ring", "fft"):  # config.rs:90-99
            raise ValueError(f"unknown autocorrelation_backend '{backend}', expected 'ring' or 'fft'")
        if cluster_update_interval is not None:
            mode = "sw" if cluster_mode is None else cluster_mode
            action = "update" if cluster_action is None else cluster_action
            if mode not in ("wolff", "sw"):
                raise ValueError(f"unknown cluster_mode '{mode}', expected 'wolff' or 'sw'")
            if action not in ("update", "observe"):
                raise ValueError(f"unknown cluster action '{action}', expected 'update' or 'observe'")
            if action == "observe" and mode == "wolff":  # config.rs:191-195
                raise ValueError("cluster_action='observe' requires cluster_mode='sw'")
            if action == "observe":
                raise ValueError("cluster_action='observe' is not implemented on the GPU sweep path")
        if collect_cluster_stats:
            raise ValueError("collect_cluster_stats is not implemented on the GPU sweep path")
        oc_wolff = True
        if overlap_cluster_update_interval is not None:  # src/lib.rs:249-262
            build = "houdayer" if overlap_cluster_build_mode is None else overlap_cluster_build_mode
            oc_mode = "wolff" if overlap_cluster_mode is None else overlap_cluster_mode
            oc_action = "update" if overlap_cluster_action is None else overlap_cluster_action
            if oc_mode not in ("wolff", "sw"):
                raise ValueError(f"unknown cluster_mode '{oc_mode}', expected 'wolff' or 'sw'")
            if oc_action not in ("update", "observe"):
                raise ValueError(f"unknown cluster action '{oc_action}', expected 'update' or 'observe'")
            if build.strip() not in ("houdayer", "houd2"):  # config.rs:117-131: houdN, jorg, cmr, a+b round-robin
                raise ValueError(f"overlap_cluster_build_mode '{build}' is not implemented on the GPU sweep path (only 'houdayer')")
            if oc_action == "observe":
                raise ValueError("overlap_cluster_action='observe' is not implemented on the GPU sweep path")
            if snapshot_interval is not None:
                raise ValueError("snapshot_interval is not implemented on the GPU sweep path")
            oc_wolff = oc_mode == "wolff"
        if pt_interval is not None and int(pt_interval) == 0:  # config.rs:197-199
            raise ValueError("pt_interval must be >= 1")
        if backend == "fft" and autocorrelation_max_lag is None:  # config.rs:200-206
            raise ValueError("autocorrelation_backend='fft' requires autocorrelation_max_lag")

        cfg = SampleCfg()
        cfg.n_sweeps = n_sweeps
        cfg.warmup_sweeps = _rust_round(n_sweeps * warmup)
        cfg.sweep_mode = _lib.SWEEP_MODES[sweep_mode]
        cfg.pt_interval = 0 if pt_interval is None else int(pt_interval)
        cfg.pt_schedule = _lib.PT_SCHEDULES[schedule]
        cfg.cluster_update_interval = 0 if cluster_update_interval is None else max(int(cluster_update_interval), 1)
        cfg.overlap_cluster_update_interval = (
            0 if overlap_cluster_update_interval is None else max(int(overlap_cluster_update_interval), 1))
        cfg.autocorrelation_max_lag = 0 if autocorrelation_max_lag is None else max(int(autocorrelation_max_lag), 1)
        cfg.snapshot_interval = 0 if snapshot_interval is None else max(int(snapshot_interval), 1)
        cfg.equilibration_diagnostic = int(bool(equilibration_diagnostic))
        cfg.cluster_mode = 1 if (cluster_update_interval is not None and cluster_mode == "wolff") else 0
        cfg.overlap_cluster_mode = 1 if oc_wolff else 0
        cfg.exact_log = int(bool(exact_log))

        T, R, D, N = self.n_temps, self.n_replicas, self.n_realizations, self.n_spins
        out = {}
        res = Results()
        for k in _F64_KEYS[:5]:
            out[k] = np.zeros(T, dtype=np.float64)
        if R >= 2:
            for k in _F64_KEYS[5:]:
                out[k] = np.zeros(T, dtype=np.float64)
            hist = np.zeros((T, N + 1), dtype=np.uint64)
            out["ql_at_q_sum"] = np.zeros((T, N + 1), dtype=np.float64)
            out["ql2_at_q_sum"] = np.zeros((T, N + 1), dtype=np.float64)
            res.overlap_histogram = hist.ctypes.data_as(_lib._PU64)
            res.ql_at_q_sum = out["ql_at_q_sum"].ctypes.data_as(_lib._PD)
            res.ql2_at_q_sum = out["ql2_at_q_sum"].ctypes.data_as(_lib._PD)
            if D > 1 and per_sample:  # src/lib.rs:385-411
                out["per_sample_overlap_histogram"] = np.zeros((D, T, N + 1), dtype=np.uint64)
                out["per_sample_ql_at_q_sum"] = np.zeros((D, T, N + 1), dtype=np.float64)
                out["per_sample_ql2_at_q_sum"] = np.zeros((D, T, N + 1), dtype=np.float64)
                res.per_sample_overlap_histogram = out["per_sample_overlap_histogram"].ctypes.data_as(_lib._PU64)
                res.per_sample_ql_at_q_sum = out["per_sample_ql_at_q_sum"].ctypes.data_as(_lib._PD)
                res.per_sample_ql2_at_q_sum = out["per_sample_ql2_at_q_sum"].ctypes.data_as(_lib._PD)
        for k in out:
            if k in _F64_KEYS:
                setattr(res, k, out[k].ctypes.data_as(_lib._PD))
        means = np.zeros((D, 11, T), dtype=np.float64)
        res.per_sample_means = means.ctypes.data_as(_lib._PD)
        pt = None
        if pt_interval is not None:  # src/lib.rs:458-490
            pt = {
                "edge_attempts": np.zeros((D, max(T - 1, 0)), dtype=np.uint64),
                "edge_acceptances": np.zeros((D, max(T - 1, 0)), dtype=np.uint64),
                "round_trips": np.zeros((D, R, T), dtype=np.uint64),
            }
            res.pt_edge_attempts = pt["edge_attempts"].ctypes.data_as(_lib._PU64)
            res.pt_edge_acceptances = pt["edge_acceptances"].ctypes.data_as(_lib._PU64)
            res.pt_round_trips = pt["round_trips"].ctypes.data_as(_lib._PU64)

        taus = None
        if autocorrelation_max_lag is not None:  # src/lib.rs:545-556 (ring accumulation for both backends)
            out["mags2_tau"] = np.zeros(T, dtype=np.float64)
            res.mags2_tau = out["mags2_tau"].ctypes.data_as(_lib._PD)
            if R >= 2:
                out["overlap2_tau"] = np.zeros(T, dtype=np.float64)
                res.overlap2_tau = out["overlap2_tau"].ctypes.data_as(_lib._PD)
            taus = np.zeros((D, 2, T), dtype=np.float64)
            res.per_sample_taus = taus.ctypes.data_as(_lib._PD)

        equil = None
        if equilibration_diagnostic:  # src/lib.rs:559-574
            buf = (C.c_int64 * 80)()
            n_ckpt = self._lib.pp_equil_checkpoints(n_sweeps, buf)
            out["equil_sweeps"] = np.asarray([buf[i] for i in range(n_ckpt)], dtype=np.uint64)
            out["equil_energy_avg"] = np.zeros((n_ckpt, T), dtype=np.float64)
            out["equil_link_overlap_avg"] = np.zeros((n_ckpt, T), dtype=np.float64)
            res.equil_energy_avg = out["equil_energy_avg"].ctypes.data_as(_lib._PD)
            res.equil_link_overlap_avg = out["equil_link_overlap_avg"].ctypes.data_as(_lib._PD)
            equil = np.zeros((D, n_ckpt, 2, T), dtype=np.float64)
            res.per_sample_equil = equil.ctypes.data_as(_lib._PD)

        cb = _lib.ON_SWEEP(lambda _user, sweep: on_sweep(int(sweep))) if on_sweep is not None else None
        flag_ptr = None
        if interrupt is not None:  # an int32 numpy scalar array the caller may set to non-zero
            flag_ptr = interrupt.ctypes.data
        if profile:  # measurement hook: CUDA events around every sweep-kernel launch of this call (bench.py's roofline pass)
            _lib.check(self._lib.pp_debug_set_profile(self._h, 1))
        status = self._lib.pp_sample(self._h, C.byref(cfg), C.byref(res), flag_ptr,
                                     C.cast(cb, C.c_void_p) if cb is not None else None, None)
        _lib.check(status)
        tm = _lib.Timing()  # measurement hook (pp_debug_last_timing): not part of the reference boundary
        _lib.check(self._lib.pp_debug_last_timing(self._h, C.byref(tm)))
        self.last_sweep_loop_ms = float(tm.sweep_loop_ms)
        self.last_kernel_launches = int(tm.kernel_launches)
        self.last_sweep_kernel_ms = float(tm.sweep_kernel_ms)
        self.last_sweep_kernel_launches = int(tm.sweep_kernel_launches)
        self.last_per_sample_means = means
        self.last_per_sample_taus = taus
        self.last_per_sample_equil = equil
        if R >= 2:
            out["overlap_histogram"] = [hist[t].copy() for t in range(T)]  # list of u64[N+1], src/lib.rs:358-366
        if pt is not None:
            out["per_disorder"] = {"parallel_tempering": pt}
        return out

    # ------------------------------------------------------------------ src/lib.rs:620-633
    def get_spins(self, realization=0):
        out = np.zeros(self.n_replicas * self.n_temps * self.n_local_spins, dtype=np.int8)
        _lib.check(self._lib.pp_get_spins(self._h, int(realization), out.ctypes.data))
        return out

    def reset(self, seed=None):
        _lib.check(self._lib.pp_reset(self._h, int(seed is not None), 0 if seed is None else int(seed)))

    # ------------------------------------------------------------------ state / operator access (parity tests)
    def get_system_ids(self, realization=0):
        out = np.zeros(self.n_replicas * self.n_temps, dtype=np.int64)
        _lib.check(self._lib.pp_get_system_ids(self._h, int(realization), out.ctypes.data))
        return out

    def get_energies(self, realization=0):
        out = np.zeros(self.n_replicas * self.n_temps, dtype=np.float32)
        _lib.check(self._lib.pp_get_energies(self._h, int(realization), out.ctypes.data))
        return out

    def set_spins(self, spins, realization=0):
        a = np.ascontiguousarray(spins, dtype=np.int8).reshape(-1)
        if a.size != self.n_replicas * self.n_temps * self.n_local_spins:
            raise ValueError("spins must have n_systems * n_spins entries")
        _lib.check(self._lib.pp_set_spins(self._h, int(realization), a.ctypes.data))

    def set_system_ids(self, ids, realization=0):
        a = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        if a.size != self.n_replicas * self.n_temps:
            raise ValueError("system_ids must have n_systems entries")
        _lib.check(self._lib.pp_set_system_ids(self._h, int(realization), a.ctypes.data))

    def op_sweep(self, sweep_mode, sweep_index, exact_log=True):
        _lib.check(self._lib.pp_op_sweep(self._h, _lib.SWEEP_MODES[sweep_mode], int(sweep_index), int(exact_log)))

    def op_energies_mags(self):
        S = self.n_replicas * self.n_temps
        e = np.zeros((self.n_realizations, S), dtype=np.float32)
        m = np.zeros((self.n_realizations, S), dtype=np.int64)
        _lib.check(self._lib.pp_op_energies_mags(self._h, e.ctypes.data, m.ctypes.data))
        return e, m

    def op_overlap(self):
        shape = (self.n_realizations, self.n_replicas // 2, self.n_temps)
        ds = np.zeros(shape, dtype=np.int64)
        dl = np.zeros(shape, dtype=np.int64)
        _lib.check(self._lib.pp_op_overlap(self._h, ds.ctypes.data, dl.ctypes.data))
        return ds, dl

    def op_pt(self, pt_schedule, pt_event):
        _lib.check(self._lib.pp_op_pt(self._h, _lib.PT_SCHEDULES[pt_schedule], int(pt_event)))


def nccl_comm_cached(device, world, rank) -> bool:
    """True when this process already holds the engine's communicator of (device, world, rank): no bootstrap token is needed."""
    return bool(_lib.load().pp_nccl_comm_cached(int(device), int(world), int(rank)))


def nccl_unique_id() -> bytes:
    """Bootstrap token for a slab-decomposed lattice: call on rank 0, broadcast to the other ranks."""
    lib = _lib.load()
    buf = np.zeros(_lib.NCCL_ID_BYTES, dtype=np.uint8)
    _lib.check(lib.pp_nccl_unique_id(buf.ctypes.data))
    return buf.tobytes()


def colouring(lattice_shape, neighbor_offsets=None):
    """The checkerboard visit order (RNG-SPEC): colour of every site, host-only."""
    lib = _lib.load()
    shape = np.asarray(lattice_shape, dtype=np.int64)
    offsets = None if neighbor_offsets is None else np.ascontiguousarray(neighbor_offsets, dtype=np.int64)
    colour = np.zeros(int(np.prod(shape)), dtype=np.uint16)
    n = C.c_int32(0)
    _lib.check(lib.pp_colouring(len(shape), shape.ctypes.data, 0 if offsets is None else len(offsets),
                                None if offsets is None else offsets.ctypes.data, colour.ctypes.data, C.byref(n)))
    return colour, n.value


def metropolis_lookup(temperatures, n_neighbors, sweep_mode="metropolis"):
    lib = _lib.load()
    t = np.ascontiguousarray(temperatures, dtype=np.float32)
    table = np.zeros((len(t), 4 * n_neighbors + 1), dtype=np.uint32)
    _lib.check(lib.pp_metropolis_lookup(t.ctypes.data, len(t), n_neighbors, _lib.SWEEP_MODES[sweep_mode], table.ctypes.data))
    return table
