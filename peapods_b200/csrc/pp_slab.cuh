// pp_slab.cuh — host driver of the slab-decomposed single-lattice path (kernels: pp_kernels_slab.cuh).
//
// One lattice too large for tables (or for one GPU) is cut along x0 into `ranks` slabs of P = L0 / ranks planes,
// one slab per process / GPU (torchrun, one rank per GPU).  Per colour half-step:
// byte storage:
//   main stream:  boundary planes (1 and P)  ->  interior planes (2..P-1)            -> wait for the halos
//   comm stream:                  wait boundary -> send plane 1 down, plane P up; receive the two halo planes
// bit-packed storage (round 2): the boundary launch no longer sits in front of the interior launch on one stream —
//   comm stream:  wait for the previous half-step's interior -> boundary planes -> send / receive the UPDATED COLOUR of the planes
//   main stream:  interior planes (they only read what the previous half-step left behind) -> wait for the halos
// so the NCCL transfer of the just-updated boundary planes overlaps the interior update.  NCCL is loaded with
// dlopen at first use (the library itself does not link against it); `rank = -1` keeps all slabs in this process on
// one device and moves halos with device copies — the same sequencing, used by the single-GPU parity tests.
// The reference has no counterpart (single address space); energies are integer sums combined with
// ncclAllReduce, so every rank sees identical energies and replays identical parallel-tempering decisions.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "pp_kernels_slab.cuh"
#include "pp_kernels_slabp.cuh"

namespace pp {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // a copy already loaded by the host process (e.g. torch's bundled NCCL) is reused by soname
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        api.error = std::string("libnccl.so.2 not found: ") + dlerror();
        return api;
    }
#define PP_NCCL_SYM(field, sym)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));     \
    if (!api.field) api.error = std::string("NCCL symbol missing: ") + sym;
    PP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PP_NCCL_SYM(GroupStart, "ncclGroupStart")
    PP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PP_NCCL_SYM(Send, "ncclSend")
    PP_NCCL_SYM(Recv, "ncclRecv")
    PP_NCCL_SYM(AllReduce, "ncclAllReduce")
    PP_NCCL_SYM(AllGather, "ncclAllGather")
    PP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PP_NCCL_SYM
    return api;
}

struct SlabState {
    int ranks = 1, rank = 0;           // rank = -1: every slab lives in this process (emulation)
    int L0 = 0, L1 = 0, L2 = 0, P = 0;  // P own planes per slab
    int64_t plane = 0;
    std::vector<SlabView> parts;       // local slabs: one, or `ranks` when emulated (byte storage)
    std::vector<uint8_t *> buffers;
    // bit-packed storage (pp_kernels_slabp.cuh), chosen when shape[2] is a multiple of 64: pparts instead of parts
    bool packed = false;
    std::vector<SlabPView> pparts;
    int nm_metro = 7, nm_gibbs = 7;    // thresholds the packed kernel compares per site (3: Metropolis fast path)
    bool mono_metro = true, mono_gibbs = true;  // acceptance counts grow with the number of unsatisfied bonds
    bool comm_cached = false;          // the communicator belongs to the process-wide cache (never destroyed here)
    cudaEvent_t ev_main = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
    unsigned long long *d_partial = nullptr;  // [2 * S]
    int64_t local_planes() const { return (int64_t)(packed ? pparts.size() : parts.size()) * P; }
};

}  // namespace pp
