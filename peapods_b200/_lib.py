"""ctypes view of the C ABI declared in ``include/peapods_b200.h``.

There is no CPU fallback: if the shared library is missing or cannot be loaded, importing the
engine fails loudly."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libpeapods_b200.so"

PP_OK, PP_ERR_INVALID, PP_ERR_UNSUPPORTED, PP_ERR_INTERRUPTED, PP_ERR_CUDA, PP_ERR_NCCL, PP_ERR_OOM = range(7)
SWEEP_MODES = {"metropolis": 0, "gibbs": 1}
PT_SCHEDULES = {"single_random_edge": 0, "full_ladder": 1}
LAYOUTS = {"auto": 0, "int8": 1, "msc": 2, "slab": 3}
NCCL_ID_BYTES = 128
LAYOUT_NAMES = {v: k for k, v in LAYOUTS.items()}

_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_dims", C.c_int32),
        ("shape", C.c_void_p),
        ("n_offsets", C.c_int32),
        ("offsets", C.c_void_p),
        ("coupling_kind", C.c_int32),
        ("couplings", C.c_void_p),
        ("n_disorder", C.c_int64),
        ("sample_offset", C.c_int64),
        ("temperatures", C.c_void_p),
        ("n_temps", C.c_int32),
        ("n_replicas", C.c_int32),
        ("seed", C.c_uint64),
        ("layout", C.c_int32),
        ("device", C.c_int32),
        ("slab_ranks", C.c_int32),
        ("slab_rank", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
        ("system_ranks", C.c_int32),
        ("system_rank", C.c_int32),
    ]


class Timing(C.Structure):
    _fields_ = [("sweep_loop_ms", C.c_double), ("kernel_launches", C.c_int64), ("sweep_kernel_ms", C.c_double),
                ("sweep_kernel_launches", C.c_int64)]


class SampleCfg(C.Structure):
    _fields_ = [
        ("n_sweeps", C.c_int64),
        ("warmup_sweeps", C.c_int64),
        ("sweep_mode", C.c_int32),
        ("pt_interval", C.c_int64),
        ("pt_schedule", C.c_int32),
        ("cluster_update_interval", C.c_int64),
        ("overlap_cluster_update_interval", C.c_int64),
        ("autocorrelation_max_lag", C.c_int64),
        ("snapshot_interval", C.c_int64),
        ("equilibration_diagnostic", C.c_int32),
        ("exact_log", C.c_int32),
        ("cluster_mode", C.c_int32),
        ("overlap_cluster_mode", C.c_int32),
    ]


RESULT_F64 = ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4",
              "link_overlap", "link_overlap2", "link_overlap4")


class Results(C.Structure):
    _fields_ = (
        [(n, _PD) for n in RESULT_F64]
        + [("overlap_histogram", _PU64), ("ql_at_q_sum", _PD), ("ql2_at_q_sum", _PD)]
        + [("per_sample_overlap_histogram", _PU64), ("per_sample_ql_at_q_sum", _PD), ("per_sample_ql2_at_q_sum", _PD)]
        + [("pt_edge_attempts", _PU64), ("pt_edge_acceptances", _PU64), ("pt_round_trips", _PU64)]
        + [("per_sample_means", _PD)]
        + [("mags2_tau", _PD), ("overlap2_tau", _PD), ("per_sample_taus", _PD)]
        + [("equil_energy_avg", _PD), ("equil_link_overlap_avg", _PD), ("per_sample_equil", _PD)]
    )


ON_SWEEP = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64)

# every symbol include/peapods_b200.h declares
SIGNATURES = {
    "pp_last_error": (C.c_char_p, []),
    "pp_abi_version": (C.c_int32, []),
    "pp_struct_size": (C.c_int64, [C.c_int32]),
    "pp_equil_checkpoints": (C.c_int32, [C.c_int64, C.POINTER(C.c_int64)]),
    "pp_colouring": (C.c_int32, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "pp_metropolis_lookup": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pp_realization_seed": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "pp_create": (C.c_int32, [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "pp_destroy": (None, [C.c_void_p]),
    "pp_sample": (C.c_int32, [C.c_void_p, C.POINTER(SampleCfg), C.POINTER(Results), C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_reset": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint64]),
    "pp_get_spins": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_get_system_ids": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_get_energies": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_get_layout": (C.c_int32, [C.c_void_p]),
    "pp_local_spin_count": (C.c_int64, [C.c_void_p]),
    "pp_nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "pp_nccl_comm_cached": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "pp_uses_msc3d": (C.c_int32, [C.c_void_p]),
    "pp_debug_set_profile": (C.c_int32, [C.c_void_p, C.c_int32]),
    "pp_debug_last_timing": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "pp_slab_packed": (C.c_int32, [C.c_void_p]),
    "pp_rows_packed": (C.c_int32, [C.c_void_p]),
    "pp_sys_words": (C.c_int32, [C.c_void_p]),
    "pp_set_spins": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_set_system_ids": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p]),
    "pp_op_sweep": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32]),
    "pp_op_energies_mags": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_op_overlap": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_slice_sweep": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_slice_energies_mags": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_slice_overlap": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pp_slice_pt": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pp_op_pt": (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32]),
}

_lib = None


def load():
    """Load the CUDA extension; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m peapods_b200.build, or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.pp_abi_version() != 4:
        raise ImportError("libpeapods_b200.so ABI version mismatch")
    for which, struct in enumerate((ModelDesc, SampleCfg, Results)):
        if lib.pp_struct_size(which) != C.sizeof(struct):
            raise ImportError(f"libpeapods_b200.so: layout of {struct.__name__} differs from include/peapods_b200.h "
                              f"({lib.pp_struct_size(which)} vs {C.sizeof(struct)} bytes)")
    _lib = lib
    return lib


def check(status: int):
    if status == PP_OK:
        return
    msg = load().pp_last_error().decode()
    if status == PP_ERR_INTERRUPTED:
        raise KeyboardInterrupt(msg)  # src/lib.rs:327-333
    if status in (PP_ERR_INVALID, PP_ERR_UNSUPPORTED):
        raise ValueError(msg)
    if status == PP_ERR_OOM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
