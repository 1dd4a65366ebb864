// pp_device.cuh — shared device-side definitions (model view, reductions).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pp_rng.cuh"

namespace pp {

enum CouplingClass : int {
    COUP_FERRO = 0,  // all +1, nothing stored
    COUP_UNIT = 1,   // every coupling in {-1,0,+1} and every T eligible (mcmc/sweep.rs:109-118): int8 / sign words
    COUP_F32 = 2     // anything else: fp32 couplings, log-form thresholds
};

// Plain-old-data view of the model handed to every kernel by value.
struct ModelView {
    int64_t N;          // spins per system
    int z;              // forward directions
    int T, R, S, P;     // temps, replicas, systems per realization, replica pairs
    int64_t D;          // realizations in this handle
    int64_t sample_offset;
    int sys_lo, sys_hi; // systems this process updates (system-split handles: pp_model_desc.system_ranks); [0, S) otherwise
    int n_colours;
    int coupling_class;
    uint64_t seed;      // current dynamics root seed
    // geometry tables
    const uint32_t *nbr;          // [N][2z] fwd/bwd interleaved
    const uint32_t *order;        // [N] colour-sorted sites
    const uint32_t *colour_start; // [n_colours+1]
    const uint32_t *perm;         // [N] site -> storage index of the multispin words (nullptr: identity); when set,
                                  // nbr / order / Jw are expressed in storage space
    // couplings
    const int8_t *J8;   // [D][N][z]   (COUP_UNIT, int8 layout)
    const float *Jf;    // [D][N][z]   (COUP_F32)
    const uint32_t *Jw; // [G][z][N]   (MSC sign words, bit=1: J=-1)
    // state
    int8_t *spins;      // [D][S][N] system-major (int8 layout)
    uint32_t *words;    // [G][S][N] slot-major   (MSC layout)
    int32_t *system_ids;// [D][S] slot -> system
    float *energies;    // [D][S] by system
    long long *mags;    // [D][S] by system
    const float *temps; // [T]
    const uint32_t *lut;// [T][4z+1] acceptance counts (Metropolis or Gibbs)
    const float *logtab;// [2^24] host-libm logf(d/2^24) or nullptr
    const float *glogtab;// [2^24] host-libm logf(u/(1-u)) or nullptr
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in thread 0
template <typename Tv>
__device__ __forceinline__ Tv block_sum(Tv v, Tv *scratch /* [32] */) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    Tv r = 0;
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? scratch[lane] : Tv(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
    }
    __syncthreads();
    return r;
}

}  // namespace pp
