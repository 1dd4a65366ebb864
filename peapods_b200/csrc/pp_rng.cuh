// pp_rng.cuh — RNG-SPEC v2: counter-based Philox4x32-7 keyed by (seed, sweep, site-rank, stream).
//
// v2 (round 2): 7 rounds instead of 10 — the smallest round count Salmon et al. (SC'11, Table 2) report as Crush-resistant,
// shipped by Random123 as philox4x32_R(7, ...); the generator's 32x32->64 multiplies (IMAD.WIDE: 0.17 per cycle per SM
// sub-partition on B200, profiles/r1_ubench_pipes.log) are the largest single cost of every sweep kernel.  Both round counts are
// pinned by the published Random123 known-answer vectors (see tests/).  v2 also adds the packed draw mapping
// of the bit-packed single-lattice kernels (pp_kernels_slabp.cuh: 32 ranks per six calls, every generated bit used).
//
// Replaces the per-system xoshiro256** streams of the reference
// (spin-sim/src/simulation/realization.rs:168-175, spin-sim/src/parallel.rs:27-33): a draw is a pure
// function of (key, counter), so any thread can produce the draw of any site without shared state.
//
//   key      = 64-bit per-realization dynamics seed  splitmix64(root ^ splitmix64(r))   (src/lib.rs:30-32)
//              (MSC layout, sweep draws: one key per 32-sample word group)
//   counter  = { rank >> 2, sweep_index | pt_event, stream, tag | colour }
//   draw24   = out[rank & 3] >> 8      (the 24 high bits of one u32, as rand's Standard f32:
//                                       spin-sim/src/mcmc/sweep.rs:179-181)
//   rank     = index of the site among the sites of its colour (ascending site index), so one
//              Philox call serves four consecutive same-colour sites.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#else
#define PP_HD inline
#endif

namespace pp {

constexpr uint32_t TAG_INIT = 0x00010000u;
constexpr uint32_t TAG_SWEEP = 0x00020000u;
constexpr uint32_t TAG_PT = 0x00030000u;
constexpr uint32_t TAG_SWEEP_MSC = 0x00040000u;
constexpr uint32_t TAG_SWEEP_PACKED = 0x000A0000u;  // counter = {rank >> 5, sweep, system, tag | call << 8 | colour}, call = 0..5
constexpr uint32_t TAG_SWEEP_SYSQ = 0x000B0000u;    // counter = {colour rank, sweep, system >> 2, tag | colour}; system s draws out[s & 3] >> 8
constexpr uint64_t MSC_KEY_DOMAIN = 0x6D73635F67726F75ull;

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

struct u32x4 {
    uint32_t x, y, z, w;
};

PP_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#ifndef PP_PHILOX_ROUNDS
#define PP_PHILOX_ROUNDS 7
#endif
constexpr int PHILOX_ROUNDS = PP_PHILOX_ROUNDS;

PP_HD u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < PHILOX_ROUNDS; round++) {
        uint32_t hi0 = mulhi32(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        uint32_t hi1 = mulhi32(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return u32x4{c0, c1, c2, c3};
}

// the same generator with the round keys k + i * W precomputed once (a thread that makes several calls with one key)
struct PhiloxKeys {
    uint32_t k0[PHILOX_ROUNDS], k1[PHILOX_ROUNDS];
};
PP_HD PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys ks;
#pragma unroll
    for (int round = 0; round < PHILOX_ROUNDS; round++) {
        ks.k0[round] = k0 + (uint32_t)round * PHILOX_W0;
        ks.k1[round] = k1 + (uint32_t)round * PHILOX_W1;
    }
    return ks;
}
PP_HD u32x4 philox4x32_k(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &ks) {
#pragma unroll
    for (int round = 0; round < PHILOX_ROUNDS; round++) {
        uint32_t hi0 = mulhi32(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        uint32_t hi1 = mulhi32(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ ks.k0[round];
        uint32_t n2 = hi0 ^ c3 ^ ks.k1[round];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return u32x4{c0, c1, c2, c3};
}

PP_HD uint64_t splitmix64(uint64_t value) {  // realization.rs:9-15
    value += 0x9E3779B97F4A7C15ull;
    uint64_t mixed = value;
    mixed = (mixed ^ (mixed >> 30)) * 0xBF58476D1CE4E5B9ull;
    mixed = (mixed ^ (mixed >> 27)) * 0x94D049BB133111EBull;
    return mixed ^ (mixed >> 31);
}

PP_HD uint64_t realization_seed(uint64_t root, uint64_t r) {  // src/lib.rs:30-32
    return splitmix64(root ^ splitmix64(r));
}

PP_HD uint64_t msc_group_key(uint64_t root, uint64_t group) {
    return splitmix64(root ^ splitmix64(MSC_KEY_DOMAIN ^ group));
}

PP_HD uint32_t pick(const u32x4 &v, uint32_t lane) {
    return lane == 0 ? v.x : lane == 1 ? v.y : lane == 2 ? v.z : v.w;
}

// The four 24-bit draws of the ranks 4 r4 .. 4 r4 + 3 in the PACKED mapping (32 ranks per six calls, counter = {rank >> 5, sweep,
// stream, tag | call << 8 | colour}): for a kernel that owns four ranks at a time (the int8 row kernels of a handle whose other
// kernels are bit-packed, so that the handle's trajectory does not depend on which kernel runs a sweep).  One or two calls.
PP_HD void packed_draws4(uint32_t r4, uint32_t sweep, uint32_t stream, uint32_t tag_colour, uint32_t k0, uint32_t k1, uint32_t (&d)[4]) {
    const uint32_t block = r4 >> 3, i0 = 3u * (r4 & 7u);
    const uint32_t c0 = i0 >> 2, c1 = (i0 + 2u) >> 2;
    const u32x4 a = philox4x32(block, sweep, stream, tag_colour | (c0 << 8), k0, k1);
    u32x4 b = a;
    if (c1 != c0) b = philox4x32(block, sweep, stream, tag_colour | (c1 << 8), k0, k1);
    const uint32_t A = pick(a, i0 & 3u);
    const uint32_t B = pick(((i0 + 1u) >> 2) == c0 ? a : b, (i0 + 1u) & 3u);
    const uint32_t C = pick(b, (i0 + 2u) & 3u);
    d[0] = A >> 8; d[1] = B >> 8; d[2] = C >> 8;
    d[3] = ((A & 255u) << 16) | ((B & 255u) << 8) | (C & 255u);
}

}  // namespace pp
