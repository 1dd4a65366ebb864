// pp_kernels_slabp.cuh — the bit-packed form of the slab layout: one large 3-D hypercubic ferromagnet with ONE BIT per
// spin, 32 same-colour sites of a row per u32 word, stride geometry (no tables), slab-decomposed along x0.
//
// Replaces, for this layout,
//   metropolis_sweep / gibbs_sweep (lookup rule)     spin-sim/src/mcmc/sweep.rs:170-185, 220-284
//   compute_energies_and_magnetizations_into         spin-sim/src/spins/energy.rs:59-110
//   Realization::new / reset (spin draw)             spin-sim/src/simulation/realization.rs:177-182
//
// Storage: u32 [S][P + 2][2][L1][W], W = L2 / 64.  Plane p (1..P own planes, 0 and P + 1 halos; global x0 = first_plane + p - 1)
// holds its two checkerboard colours h = (x0 + x1 + x2) & 1 as two halves of L1 * W words; bit b of word w of row x1 in half h is
// the site x2 = 64 w + 2 b + off with off = (h + x0 + x1) & 1, bit = 1: spin -1.  The +-x0 and +-x1 neighbours of the 32 sites
// of a word are the word at the same position of the other half of the neighbouring plane / row; the two x2 neighbours are
// the other half's word at the same position and that word shifted by one bit (completed by the adjacent word).  A ferromagnetic
// bond is unsatisfied iff the two spins differ, so the six bond words are six XORs, two bit-sliced full adders count them, and
// the acceptance rule  flip <=> draw < table[t][2 * unsat]  (sweep.rs:182-184) is evaluated for 32 sites at once from
// per-site threshold masks: mask_u bit b = [draw_b < table[t][2u]].  The thresholds grow with unsat (checked on the host), so
// with L_b = #{u : draw_b < table[t][2u]}  the rule is  unsat_b + L_b >= 7  — one short carry chain per word.
//
// Draws: RNG-SPEC v2 packed mapping (pp_rng.cuh): the 32 sites of a word are the ranks 32 q .. 32 q + 31 of their colour class
// (rank = index among the sites of the colour in ascending site order), and share six Philox calls, every generated bit used:
// 2.6 IMAD.WIDE per site instead of the 5 of the one-call-per-four-sites mapping, no per-site byte extraction.
//
// HBM: one bit read + one bit written per attempt (0.25 B; SURVEY.md 8d "C5-packed"), neighbour words through L1 / L2.  The
// kernel is bound by the generator's multiplies and the 3 compares per site, not by memory.
#pragma once
#include "pp_device.cuh"

namespace pp {

struct SlabPView {
    uint32_t *words;       // [S][P + 2][2][L1][W]
    int P, L1, W;          // own planes, rows, words per row and colour (L2 / 64)
    int64_t half;          // L1 * W: one colour of one plane
    int64_t plane;         // 2 * half
    int64_t sys_stride;    // (P + 2) * plane
    int64_t first_plane;   // global x0 of local plane 1
    uint32_t k0, k1;       // Philox key of the realization
};

constexpr int SLABP_THREADS = 256;

__device__ __forceinline__ uint32_t slabp_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t slabp_maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

// One colour half-step of the local planes [pa, pa + np) and, when pb > 0, of plane pb as one more grid row (the two boundary
// planes go in one launch).  grid = (ceil(half / 256), np [+ 1], slots); one thread = one word = 32 attempts.
//   NM      thresholds compared per site: 3 = Metropolis on 6 neighbours (unsat >= 3 always flips, the three others are
//           < 2^24: host-checked), 7 = any table that grows with unsat (heat bath)
//   UPDATE  false: no draws, no store (energy / magnetisation of the current state through the same code)
//   ACC     add the unsatisfied bonds of the pass's sites AFTER the update and the down spins of both colours to
//           partial[2 sys], partial[2 sys + 1]: on the bipartite lattice the six bonds of the colour-1 sites are all bonds, each
//           once (energy.rs:99-108), so the colour-1 pass of a sweep yields its energy without another pass
template <int NM, bool UPDATE, bool ACC>
__global__ void __launch_bounds__(SLABP_THREADS)
slabp_sweep_kernel(ModelView m, SlabPView v, int colour, uint32_t sweep_index, int pa, int np, int pb, unsigned long long *partial) {
    const int slot = blockIdx.z;
    const int t = slot % m.T;  // realization.rs:166: temperatures repeat with period T
    const int64_t idx = (int64_t)blockIdx.x * SLABP_THREADS + threadIdx.x;
    const bool live = idx < v.half;
    const int p = (int)blockIdx.y < np ? pa + (int)blockIdx.y : pb;
    const uint32_t sys = (uint32_t)m.system_ids[slot];  // parallel.rs:27-33: spins by system, temperature by slot
    uint32_t unsat_sum = 0, down_sum = 0;
    if (live) {
        const int x1 = (int)(idx / v.W), w = (int)(idx - (int64_t)x1 * v.W);
        uint32_t *slf = v.words + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane + (int64_t)colour * v.half;
        const uint32_t *oth = v.words + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane + (int64_t)(1 - colour) * v.half;
        const int x1m = x1 ? x1 - 1 : v.L1 - 1, x1p = x1 + 1 == v.L1 ? 0 : x1 + 1;
        const uint32_t gx0 = (uint32_t)(v.first_plane + p - 1);
        const uint32_t off = ((uint32_t)colour + gx0 + (uint32_t)x1) & 1u;  // x2 parity of this word's sites
        const uint32_t C = slf[idx], O = oth[idx];
        const uint32_t Xm = (oth - v.plane)[idx], Xp = (oth + v.plane)[idx];
        const uint32_t Ym = oth[(int64_t)x1m * v.W + w], Yp = oth[(int64_t)x1p * v.W + w];
        const uint32_t E = oth[(int64_t)x1 * v.W + (off ? (w + 1 == v.W ? 0 : w + 1) : (w ? w - 1 : v.W - 1))];
        const uint32_t Sh = off ? (O >> 1) | (E << 31) : (O << 1) | (E >> 31);
        // bond words (1 = unsatisfied) and their bit-sliced count x + 2 y, x = s1 + s2, y = c1 + c2
        const uint32_t b0 = C ^ Xm, b1 = C ^ Xp, b2 = C ^ Ym, b3 = C ^ Yp, b4 = C ^ O, b5 = C ^ Sh;
        const uint32_t s1 = slabp_xor3(b0, b1, b2), c1 = slabp_maj3(b0, b1, b2);
        const uint32_t s2 = slabp_xor3(b3, b4, b5), c2 = slabp_maj3(b3, b4, b5);
        uint32_t flip = 0u;
        if (UPDATE) {
            uint32_t T[NM];
#pragma unroll
            for (int u = 0; u < NM; u++) {
                const uint32_t cnt = m.lut[t * 13 + 2 * u];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
                T[u] = NM == 3 ? cnt << 8 : cnt;             // NM == 3: cnt < 2^24, compare the 32-bit word against cnt << 8
            }
            uint32_t M[NM];  // NM == 3: M[0], M[1] = the two bit planes of L (the number of thresholds above the draw), M[2] unused
#pragma unroll
            for (int u = 0; u < NM; u++) M[u] = 0u;
            const uint32_t q = (uint32_t)(((int64_t)gx0 * v.L1 + x1) * v.W + w);  // rank >> 5 of the word's sites
            const uint32_t tagc = TAG_SWEEP_PACKED | (uint32_t)colour;
            const PhiloxKeys ks = philox_keys(v.k0, v.k1);  // six calls, one key schedule
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t wd[12];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const u32x4 o = philox4x32_k(q, sweep_index, sys, tagc | ((uint32_t)(3 * h + c) << 8), ks);
                    wd[4 * c] = o.x; wd[4 * c + 1] = o.y; wd[4 * c + 2] = o.z; wd[4 * c + 3] = o.w;
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint32_t A = wd[3 * g], B = wd[3 * g + 1], Cw = wd[3 * g + 2];
                    // fourth field of the group: the low bytes of A, B, Cw (its own low byte is never looked at)
                    const uint32_t y3 = __byte_perm(__byte_perm(Cw, B, 0x0400), A, 0x4210);
                    const uint32_t ys[4] = {A, B, Cw, y3};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t bit = 1u << (16 * h + 4 * g + j);
                        const uint32_t y = NM == 3 ? ys[j] : ys[j] >> 8;
                        if (NM == 3) {  // T[0] <= T[1] <= T[2] (host-checked): two compares place the draw, L = 2 p1 + p2
                            const bool p1 = y < T[1];
                            const bool p2 = y < (p1 ? T[0] : T[2]);
                            if (p1) M[1] |= bit;
                            if (p2) M[0] |= bit;
                        } else {
#pragma unroll
                            for (int u = 0; u < NM; u++)
                                if (y < T[u]) M[u] |= bit;
                        }
                    }
                }
            }
            const uint32_t kk = s1 & s2, oo = s1 | s2;
            if (NM == 3) {
                // L = L0 + 2 L1 thresholds lie above the draw: flip <=> x + 2 y + L >= 3, i.e. with a = x + L0, b = y + L1:
                // b >= 2, or b = 1 and a >= 1, or b = 0 and a = 3 (x = 2 and L0)
                const uint32_t L0 = M[0], L1 = M[1], y0 = c1 ^ c2, y1 = c1 & c2;
                flip = y1 | (y0 & L1) | ((y0 ^ L1) & ~y1 & (oo | L0)) | (~(c1 | c2 | L1) & kk & L0);
            } else {
                // unsat = x0 + 2 y1 + 4 y2, L = L0 + 2 L1 + 4 L2 from the nested masks M0 <= ... <= M6: flip <=> unsat + L + 1 >= 8
                const uint32_t x0 = s1 ^ s2, y1 = slabp_xor3(c1, c2, kk), y2 = slabp_maj3(c1, c2, kk);
                const uint32_t L0 = slabp_xor3(M[0], M[1], M[2]) ^ slabp_xor3(M[3], M[4], M[5]) ^ M[6];
                const uint32_t L1 = (M[5] & ~M[3]) | M[1], L2 = M[3];
                const uint32_t ca = x0 | L0, cb = slabp_maj3(y1, L1, ca);
                flip = slabp_maj3(y2, L2, cb);
            }
            slf[idx] = C ^ flip;
        }
        if (ACC) {  // a flip inverts the six bond words of its site: sums and carries of both adders are XORed with the flip mask
            unsat_sum = __popc(s1 ^ flip) + __popc(s2 ^ flip) + 2u * (__popc(c1 ^ flip) + __popc(c2 ^ flip));
            down_sum = __popc(C ^ flip) + __popc(O);
        }
    }
    if (ACC) {
        __shared__ uint32_t sh[32];
        const uint32_t tu = block_sum<uint32_t>(unsat_sum, sh);
        const uint32_t td = block_sum<uint32_t>(down_sum, sh);
        if (threadIdx.x == 0) {
            atomicAdd(&partial[2 * sys], (unsigned long long)tu);
            atomicAdd(&partial[2 * sys + 1], (unsigned long long)td);
        }
    }
}

// K0: spin -1 iff the INIT-domain draw of the SITE < 2^23 (realization.rs:180; the same site-indexed draws as every other
// layout, so a fresh handle holds the same configuration whatever its storage).  One thread = 64 consecutive sites of a row
// = one word of each colour; grid = (ceil(half / 256), P, S).
__global__ void __launch_bounds__(SLABP_THREADS) slabp_init_kernel(ModelView m, SlabPView v) {
    const int64_t idx = (int64_t)blockIdx.x * SLABP_THREADS + threadIdx.x;
    if (idx >= v.half) return;
    const int p = blockIdx.y + 1;
    const uint32_t sys = blockIdx.z;
    const int x1 = (int)(idx / v.W), w = (int)(idx - (int64_t)x1 * v.W);
    const uint32_t gx0 = (uint32_t)(v.first_plane + p - 1);
    const uint64_t site0 = (((uint64_t)gx0 * v.L1 + x1) * v.W + w) * 64ull;
    uint32_t even = 0u, odd = 0u;  // sites at even / odd x2
#pragma unroll 4
    for (int c = 0; c < 16; c++) {
        const u32x4 o = philox4x32((uint32_t)(site0 / 4 + c), 0u, sys, TAG_INIT, v.k0, v.k1);
        if ((o.x >> 8) < (1u << 23)) even |= 1u << (2 * c);
        if ((o.y >> 8) < (1u << 23)) odd |= 1u << (2 * c);
        if ((o.z >> 8) < (1u << 23)) even |= 1u << (2 * c + 1);
        if ((o.w >> 8) < (1u << 23)) odd |= 1u << (2 * c + 1);
    }
    const uint32_t ce = (gx0 + (uint32_t)x1) & 1u;  // colour of the row's even-x2 sites
    uint32_t *pl = v.words + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane;
    pl[(int64_t)ce * v.half + idx] = even;
    pl[(int64_t)(1 - ce) * v.half + idx] = odd;
}

// own planes <-> +-1 int8 in the reference's order [S][planes][L1][L2]; dir 0: unpack to ext, 1: pack from ext.
// One thread = 64 consecutive sites of a row; grid = (ceil(P * half / 256), S).
__global__ void __launch_bounds__(SLABP_THREADS)
slabp_convert_kernel(SlabPView v, int8_t *ext, int64_t ext_sys_stride, int64_t ext_off, int dir) {
    const int64_t gid = (int64_t)blockIdx.x * SLABP_THREADS + threadIdx.x;
    if (gid >= (int64_t)v.P * v.half) return;
    const uint32_t sys = blockIdx.y;
    const int p = (int)(gid / v.half) + 1;
    const int64_t idx = gid - (int64_t)(p - 1) * v.half;
    const int x1 = (int)(idx / v.W);
    const uint32_t gx0 = (uint32_t)(v.first_plane + p - 1);
    const uint32_t ce = (gx0 + (uint32_t)x1) & 1u;
    uint32_t *pl = v.words + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane;
    int8_t *e = ext + (int64_t)sys * ext_sys_stride + ext_off + gid * 64;
    if (dir == 0) {
        const uint32_t even = pl[(int64_t)ce * v.half + idx], odd = pl[(int64_t)(1 - ce) * v.half + idx];
        for (int b = 0; b < 32; b++) {
            e[2 * b] = (even >> b) & 1u ? (int8_t)-1 : (int8_t)1;
            e[2 * b + 1] = (odd >> b) & 1u ? (int8_t)-1 : (int8_t)1;
        }
    } else {
        uint32_t even = 0u, odd = 0u;
        for (int b = 0; b < 32; b++) {
            if (e[2 * b] < 0) even |= 1u << b;
            if (e[2 * b + 1] < 0) odd |= 1u << b;
        }
        pl[(int64_t)ce * v.half + idx] = even;
        pl[(int64_t)(1 - ce) * v.half + idx] = odd;
    }
}

}  // namespace pp
