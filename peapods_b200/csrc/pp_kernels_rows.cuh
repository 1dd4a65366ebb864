// pp_kernels_rows.cuh — int8 layout without per-site tables: geometry lowered to per-ROW stride tables.
//
// The table-driven kernels (pp_kernels_int8.cuh) read 2z' neighbour indices per attempt and touch spins one byte at a
// time; ncu shows them bound by the L1 load pipe at 150-230 instructions per attempt.  Whenever the colouring is
// c(x) = (sum_d a_d x_d) mod m with a_last = m/2 (hypercubic checkerboard, the 4-colour triangular colouring, ...)
// and the last extent is a multiple of 8, the colour classes alternate along a row, so
//   * a thread owns the 8 consecutive sites of one row segment: its four active sites are the bytes of one parity,
//     their colour ranks are 4 consecutive numbers (= one Philox call, RNG-SPEC), rank >> 2 = row_ordinal * L/8 + segment;
//   * the neighbour in direction k of all 8 sites is one (possibly shifted) 8-byte load from row nbr_row[row][k]:
//     per-row tables of O(rows * z') entries replace the O(N * z') site tables (lattice.rs:63-82);
//   * ferromagnets count down-spin neighbours of all 8 sites with 2z' 64-bit adds (SWAR); fp32 / +-1 couplings walk the
//     four active sites in the reference's order (sweep.rs:8-19), a thread reusing its coupling registers over
//     ROWS_NS systems of the same realization.
// Replaces the same reference functions as pp_kernels_int8.cuh; results are bit-identical to those kernels.
#pragma once
#include <string>
#include <vector>

#include "pp_device.cuh"
#include "pp_kernels_stats.cuh"
#include "pp_plan.h"

namespace pp {

constexpr int ROWS_NS = 4;  // systems a sweep thread walks with one set of coupling registers
// (one system per thread for ferromagnets, which have no coupling registers to reuse, was measured: no gain at C3, a loss
// in the resident kernel at C1)
template <int CLASS>
__host__ __device__ constexpr int rows_ns() { return ROWS_NS; }

struct RowsView {
    int L;                       // last extent
    int kpr;                     // segments per row = L / 8
    int kpr_shift;               // log2(kpr) or -1
    int m_half;                  // m / 2: rows of class (A mod m_half) hold the colours A and A + m_half
    int64_t n_rows;
    const uint8_t *row_a;        // [rows] A = sum_{d<last} a_d x_d mod m
    const uint32_t *row_ord;     // [rows] number of earlier rows of the same class
    const uint32_t *nbr_row;     // [rows][2z] row of the forward / backward neighbour in direction k
    const int32_t *dl;           // [z] last-dimension component of offset k
    const uint32_t *class_rows;  // rows sorted by class
    const uint32_t *class_start; // [m_half + 1]
    const uint64_t *keys;        // [D] realization_seed(seed, sample_offset + d)
    int packed_draws;            // 1: draws in the packed mapping (a handle whose resident kernel is bit-packed, pp_kernels_prows.cuh)
};

struct RowsPlan {
    bool ok = false;
    int L = 0, kpr = 0, kpr_shift = -1, m_half = 0;
    int64_t n_rows = 0;
    std::vector<uint8_t> row_a;
    std::vector<uint32_t> row_ord, nbr_row, class_rows, class_start;
    std::vector<int32_t> dl;
};

// host: per-row tables from the plan's linear colouring (pp_plan.h)
inline RowsPlan rows_plan(const LatticePlan &p) {
    RowsPlan q;
    const int n = p.n_dims;
    if (!p.linear_colouring || n < 1) return q;
    const int m = p.colour_mod;
    if (m % 2 != 0 || p.colour_coef[(size_t)n - 1] != m / 2) return q;
    const int64_t L = p.shape[(size_t)n - 1];
    if (L % 8 != 0) return q;
    for (int k = 0; k < p.z; k++)
        if (std::abs((long long)p.offsets[(size_t)k * n + n - 1]) >= L) return q;
    q.L = (int)L;
    q.kpr = (int)(L / 8);
    for (int b = 0; b < 30; b++)
        if (q.kpr == (1 << b)) q.kpr_shift = b;
    q.m_half = m / 2;
    q.n_rows = p.n_spins / L;
    q.row_a.resize((size_t)q.n_rows);
    q.row_ord.resize((size_t)q.n_rows);
    q.nbr_row.resize((size_t)q.n_rows * 2 * p.z);
    q.dl.resize((size_t)p.z);
    for (int k = 0; k < p.z; k++) q.dl[(size_t)k] = (int32_t)p.offsets[(size_t)k * n + n - 1];
    std::vector<uint32_t> seen((size_t)q.m_half, 0);
    std::vector<int64_t> x((size_t)std::max(n - 1, 1), 0);
    for (int64_t r = 0; r < q.n_rows; r++) {
        int64_t rem = r, a = 0;
        for (int d = n - 2; d >= 0; d--) { x[(size_t)d] = rem % p.shape[(size_t)d]; rem /= p.shape[(size_t)d]; }
        for (int d = 0; d < n - 1; d++) a += p.colour_coef[(size_t)d] * x[(size_t)d];
        q.row_a[(size_t)r] = (uint8_t)(a % m);
        q.row_ord[(size_t)r] = seen[(size_t)(a % q.m_half)]++;
        for (int k = 0; k < p.z; k++)
            for (int sgn = 0; sgn < 2; sgn++) {
                int64_t row = 0;
                for (int d = 0; d < n - 1; d++) {
                    const int64_t c = rem_euclid(x[(size_t)d] + (sgn ? -1 : 1) * p.offsets[(size_t)k * n + d], p.shape[(size_t)d]);
                    row = row * p.shape[(size_t)d] + c;
                }
                q.nbr_row[((size_t)r * p.z + k) * 2 + sgn] = (uint32_t)row;
            }
    }
    q.class_start.assign((size_t)q.m_half + 1, 0);
    for (int64_t r = 0; r < q.n_rows; r++) q.class_start[(size_t)(q.row_a[(size_t)r] % q.m_half) + 1]++;
    for (int c = 0; c < q.m_half; c++) q.class_start[(size_t)c + 1] += q.class_start[(size_t)c];
    q.class_rows.resize((size_t)q.n_rows);
    std::vector<uint32_t> fill(q.class_start.begin(), q.class_start.end() - 1);
    for (int64_t r = 0; r < q.n_rows; r++) q.class_rows[fill[(size_t)(q.row_a[(size_t)r] % q.m_half)]++] = (uint32_t)r;
    q.ok = true;
    return q;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t rows_ld8(const int8_t *p) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    return (uint64_t)v.x | ((uint64_t)v.y << 32);
}

// bytes j = 0..7: spin at x_last = (8k + j + shift) mod L of row `row` (shift = +-dl of the direction)
__device__ __forceinline__ uint64_t rows_shifted(const int8_t *row, const int L, const int k, const int shift) {
    if (shift == 0) return rows_ld8(row + 8 * k);
    int start = 8 * k + shift;
    start -= (start >= L) ? L : 0;
    start += (start < 0) ? L : 0;
    const int c0 = start >> 3, sh = (start & 7) * 8;
    const int c1 = (c0 + 1) * 8 == L ? 0 : c0 + 1;
    const uint64_t lo = rows_ld8(row + 8 * c0);
    if (sh == 0) return lo;
    const uint64_t hi = rows_ld8(row + 8 * c1);
    return (lo >> sh) | (hi << (64 - sh));
}

__device__ __forceinline__ void rows_split(const RowsView &v, const uint32_t ci, uint32_t &ri, int &k) {
    if (v.kpr_shift >= 0) {
        ri = ci >> v.kpr_shift;
        k = (int)(ci & (uint32_t)(v.kpr - 1));
    } else {
        ri = ci / (uint32_t)v.kpr;
        k = (int)(ci - ri * (uint32_t)v.kpr);
    }
}

// One row segment (work item ci of colour `colour`) of ROWS_NS slots of realization d.  spins_d = the realization's
// [S][N] spins (global memory, or a shared-memory copy in the resident kernel); lut_sm = [T][4z + 1] acceptance counts.
// ZT > 0: the number of forward directions is the compile-time constant ZT (everything stays in registers); ZT = 0: any z' <= 16.
// EACC (two-colour lattices, last colour pass): also return, per walked slot, the post-update bond sum of the thread's active
// sites, sum_i s_i h_i (fp32 couplings: in units of 1 / escale, rounded to an integer), and the down spins of its segment.
// The active sites of the last pass touch every bond exactly once, so these add up to the energy of energy.rs:99-108
// without another pass over the spins.
template <int CLASS, int ZT, bool GIBBS, bool EACC = false>
__device__ __forceinline__ void rows_sweep_body(const ModelView &m, const RowsView &v, const uint32_t *lut_sm, int8_t *spins_d,
                                                const int64_t d, const int slot0, const uint32_t ci, const int colour,
                                                const uint32_t sweep_index, const int exact_log, const float escale = 1.0f,
                                                long long *e_part = nullptr, int *dn_part = nullptr) {
    constexpr int NS = rows_ns<CLASS>();
    constexpr int ZA = ZT > 0 ? ZT : 16;
    const int z = ZT > 0 ? ZT : m.z, width = 4 * z + 1;
    const int cls = colour % v.m_half;
    uint32_t ri;
    int k;
    rows_split(v, ci, ri, k);
    const uint32_t r = v.class_rows[v.class_start[cls] + ri];
    const int off = (int)v.row_a[r] == colour ? 0 : 1;  // active sites: x_last = 8k + 2l + off
    const uint32_t q = v.row_ord[r] * (uint32_t)v.kpr + (uint32_t)k;  // colour rank >> 2
    const uint64_t key = v.keys[d];
    const int L = v.L;
    const int64_t row_off = (int64_t)r * L;
    int64_t nrf[ZA], nrb[ZA];  // byte offsets of the forward / backward neighbour rows
    int dls[ZA];
    // the (up to two) 8-byte loads that make up the shifted neighbour segment of every direction, as byte offsets inside a
    // system and a funnel-shift amount: computed once per thread, reused by every system the thread walks
    uint32_t fo0[ZA], fo1[ZA], bo0[ZA], bo1[ZA];
    int fsh[ZA], bsh[ZA];
    auto seg_plan = [&](const int64_t nrow, const int shift, uint32_t &o0, uint32_t &o1, int &sh) {
        int start = 8 * k + shift;
        start -= (start >= L) ? L : 0;
        start += (start < 0) ? L : 0;
        const int c0 = start >> 3;
        const int c1 = (c0 + 1) * 8 == L ? 0 : c0 + 1;
        sh = (start & 7) * 8;
        o0 = (uint32_t)(nrow + 8 * c0);
        o1 = (uint32_t)(nrow + 8 * c1);
    };
#pragma unroll
    for (int kk = 0; kk < ZA; kk++) {
        if (kk < z) {
            nrf[kk] = (int64_t)v.nbr_row[((size_t)r * z + kk) * 2] * L;
            nrb[kk] = (int64_t)v.nbr_row[((size_t)r * z + kk) * 2 + 1] * L;
            dls[kk] = v.dl[kk];
            seg_plan(nrf[kk], dls[kk], fo0[kk], fo1[kk], fsh[kk]);
            seg_plan(nrb[kk], -dls[kk], bo0[kk], bo1[kk], bsh[kk]);
        }
    }
    auto seg_load = [&](const int8_t *sp, const uint32_t o0, const uint32_t o1, const int sh) {
        const uint64_t lo = rows_ld8(sp + o0);
        if (sh == 0) return lo;  // warp-uniform: the shift depends on the direction only
        const uint64_t hi = rows_ld8(sp + o1);
        return (lo >> sh) | (hi << (64 - sh));
    };
    // fp32 / +-1 couplings of the four active sites, reference order: direction-major, forward then backward
    float Jf[CLASS == COUP_F32 ? 4 : 1][CLASS == COUP_F32 ? 2 * ZA : 1];
    int Ji[CLASS == COUP_UNIT ? 4 : 1][CLASS == COUP_UNIT ? 2 * ZA : 1];
    if (CLASS != COUP_FERRO) {
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const int xl = 8 * k + 2 * l + off;
            const int64_t i = row_off + xl;
#pragma unroll
            for (int kk = 0; kk < ZA; kk++) {
                if (kk < z) {
                    int xb = xl - dls[kk];
                    xb -= (xb >= L) ? L : 0;
                    xb += (xb < 0) ? L : 0;
                    const int64_t jb = nrb[kk] + xb;
                    if (CLASS == COUP_F32) {
                        const float *J = m.Jf + (size_t)d * m.N * z;
                        Jf[l][2 * kk] = J[(size_t)i * z + kk];
                        Jf[l][2 * kk + 1] = J[(size_t)jb * z + kk];
                    } else {
                        const int8_t *J = m.J8 + (size_t)d * m.N * z;
                        Ji[l][2 * kk] = J[(size_t)i * z + kk];
                        Ji[l][2 * kk + 1] = J[(size_t)jb * z + kk];
                    }
                }
            }
        }
    }
#ifndef PP_ROWS_PF
#define PP_ROWS_PF 1
#endif
    // The walked systems are independent, but the compiler cannot move one system's loads above the previous system's store
    // (same array): PF > 1 loads the segments of PF systems together before updating them.  Measured at C4 (D = 128): PF = 1
    // 156 attempts/ns, PF = 2 131, PF = 4 125 -- the kernel is issue-bound (0.65 IPC per sub-partition at 24 % occupancy), the
    // extra registers cost more than the memory-level parallelism gains; constant couplings instead of loads: no change.
    constexpr int PF = NS < PP_ROWS_PF ? NS : PP_ROWS_PF;
    uint32_t sysv[PF];
    uint64_t Cv[PF], Fv[PF][ZA], Bv[PF][ZA];  // byte j: site j / its forward / backward neighbour in direction kk
#pragma unroll
    for (int ss = 0; ss < NS; ss++) {
        const int slot = slot0 + ss;
        if (slot >= m.S) break;
        if (ss % PF == 0) {
#pragma unroll
            for (int pp = 0; pp < PF; pp++) {
                if (slot + pp < m.S) {
                    sysv[pp] = (uint32_t)m.system_ids[d * m.S + slot + pp];  // parallel.rs:27-33
                    const int8_t *sp = spins_d + (int64_t)sysv[pp] * m.N;
                    Cv[pp] = rows_ld8(sp + row_off + 8 * k);
#pragma unroll
                    for (int kk = 0; kk < ZA; kk++) {
                        if (kk < z) {
                            Fv[pp][kk] = seg_load(sp, fo0[kk], fo1[kk], fsh[kk]);
                            Bv[pp][kk] = seg_load(sp, bo0[kk], bo1[kk], bsh[kk]);
                        }
                    }
                }
            }
        }
        long long e_sum = 0;
        const uint32_t sys = sysv[ss % PF];
        if ((int)sys < m.sys_lo || (int)sys >= m.sys_hi) continue;     // system-split handle: another process updates this system
        const int t = slot % m.T;                                     // realization.rs:166
        int8_t *s = spins_d + (int64_t)sys * m.N;
        uint32_t draws[4];
        if (v.packed_draws) {
            packed_draws4(q, sweep_index, sys, TAG_SWEEP_PACKED | (uint32_t)colour, (uint32_t)key, (uint32_t)(key >> 32), draws);
        } else {
            const u32x4 o = philox4x32(q, sweep_index, sys, TAG_SWEEP | (uint32_t)colour, (uint32_t)key, (uint32_t)(key >> 32));
            draws[0] = o.x >> 8; draws[1] = o.y >> 8; draws[2] = o.z >> 8; draws[3] = o.w >> 8;
        }
        const uint64_t C = Cv[ss % PF];
        uint64_t F[ZA], B[ZA];
#pragma unroll
        for (int kk = 0; kk < ZA; kk++) {
            if (kk < z) {
                F[kk] = Fv[ss % PF][kk];
                B[kk] = Bv[ss % PF][kk];
            }
        }
        // bring the four active sites to the even bytes: every per-site shift below is a compile-time constant
        const int osh = 8 * off;
        const uint64_t Cs = C >> osh;
        uint32_t fl[2] = {0u, 0u};  // byte 2(l&1) of fl[l>>1] = 0xFE where site l flips (+1 = 0x01 <-> -1 = 0xFF)
        if (CLASS == COUP_FERRO) {
            uint64_t down = 0;  // per byte: number of down-spin neighbours
#pragma unroll
            for (int kk = 0; kk < ZA; kk++)
                if (kk < z) down += ((F[kk] >> 1) & 0x0101010101010101ull) + ((B[kk] >> 1) & 0x0101010101010101ull);
            down >>= osh;
            const uint32_t dw[2] = {(uint32_t)down, (uint32_t)(down >> 32)}, cw[2] = {(uint32_t)Cs, (uint32_t)(Cs >> 32)};
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const int nd = (int)((dw[l >> 1] >> (16 * (l & 1))) & 0xFFu);
                const bool dn = ((cw[l >> 1] >> (16 * (l & 1))) & 0x80u) != 0;
                const int idx = dn ? 4 * z - 2 * nd : 2 * nd;  // ec + 2z' with ec = -s h, h = 2z' - 2 nd (sweep.rs:178)
                const bool flip = draws[l] < lut_sm[t * width + idx];
                if (flip) fl[l >> 1] |= 0xFEu << (16 * (l & 1));
                if (EACC) e_sum += flip ? idx - 2 * z : 2 * z - idx;  // s h after the update = ec if flipped, -ec if not
            }
        } else {
            const float half_t = __fdiv_rn(m.temps[t], 2.0f);
            const uint32_t cw[2] = {(uint32_t)Cs, (uint32_t)(Cs >> 32)};
            uint32_t fw[ZA][2], bw[ZA][2];
#pragma unroll
            for (int kk = 0; kk < ZA; kk++) {
                if (kk < z) {
                    const uint64_t f = F[kk] >> osh, b = B[kk] >> osh;
                    fw[kk][0] = (uint32_t)f; fw[kk][1] = (uint32_t)(f >> 32);
                    bw[kk][0] = (uint32_t)b; bw[kk][1] = (uint32_t)(b >> 32);
                }
            }
#pragma unroll
            for (int l = 0; l < 4; l++) {
                const int w = l >> 1, bs = 16 * (l & 1);
                const uint32_t sbyte = (cw[w] >> bs) & 0xFFu;
                const uint32_t draw = draws[l];
                float h = 0.0f;
                int hi = 0;
#pragma unroll
                for (int kk = 0; kk < ZA; kk++) {
                    if (kk < z) {
                        if (CLASS == COUP_F32) {
                            // sweep.rs:10-17: forward then backward, no fused multiply-add; a product with a +-1 spin is
                            // the coupling with its sign bit flipped, exactly what the f32 multiplication returns.  The spin's
                            // sign bit (bit bs + 7 of the word) goes straight to bit 31: one shift, one LOP3, one add per term.
                            h = __fadd_rn(h, __uint_as_float(__float_as_uint(Jf[l][2 * kk]) ^ ((fw[kk][w] << (24 - bs)) & 0x80000000u)));
                            h = __fadd_rn(h, __uint_as_float(__float_as_uint(Jf[l][2 * kk + 1]) ^ ((bw[kk][w] << (24 - bs)) & 0x80000000u)));
                        } else {
                            const uint32_t fb = (fw[kk][w] >> bs) & 0xFFu, bb = (bw[kk][w] >> bs) & 0xFFu;
                            hi += (int)(int8_t)fb * Ji[l][2 * kk] + (int)(int8_t)bb * Ji[l][2 * kk + 1];
                        }
                    }
                }
                bool flip;
                if (CLASS == COUP_F32) {
                    // eng_change = -s_i * h (sweep.rs:43-44): h with its sign flipped when s_i = +1
                    const float eng_change = __uint_as_float(__float_as_uint(h) ^ ((~cw[w] << (24 - bs)) & 0x80000000u));
                    const float u = (float)draw * (1.0f / 16777216.0f);
                    float lg;
                    if (!GIBBS)  // sweep.rs:256; production mode uses the hardware log2 (relative error ~1e-7 on the threshold)
                        lg = exact_log ? m.logtab[draw] : __logf(u);
                    else         // sweep.rs:279-282
                        lg = exact_log ? m.glogtab[draw] : __logf(__fdividef(u, __fsub_rn(1.0f, u)));
                    flip = eng_change >= __fmul_rn(half_t, lg);
                } else {
                    const int si = (int)(int8_t)sbyte;
                    flip = draw < lut_sm[t * width + (-si * hi) + 2 * z];  // sweep.rs:178-184
                }
                if (flip) fl[w] |= 0xFEu << bs;
                if (EACC) {
                    if (CLASS == COUP_F32) {
                        const float ec = __uint_as_float(__float_as_uint(h) ^ ((~sbyte & 0x80u) << 24));  // -s h before the update
                        e_sum += __float2ll_rn(__fmul_rn(flip ? ec : -ec, escale));
                    } else {
                        const int sh = (int)(int8_t)sbyte * hi;
                        e_sum += flip ? -sh : sh;
                    }
                }
            }
        }
        const uint64_t flips = ((uint64_t)fl[0] | ((uint64_t)fl[1] << 32)) << osh;
        const uint64_t out = C ^ flips;
        *reinterpret_cast<uint2 *>(s + row_off + 8 * k) = make_uint2((uint32_t)out, (uint32_t)(out >> 32));
        if (EACC) {
            e_part[ss] = e_sum;
            dn_part[ss] = __popcll(out & 0x8080808080808080ull);
        }
    }
}

// One colour class of one sweep.  grid = (D * ceil(S / NS), segment blocks); a thread owns one row segment and walks
// NS = rows_ns<CLASS>() slots of one realization.  EACC: the pass also delivers energies (+ magnetisation sums) -- warp
// sums (REDUX) meet in shared memory, block sums in acc[2 * sys + {0, 1}] (64-bit integer atomics: order-independent), and
// the block of a slot group that arrives last converts them and re-zeroes the scratch (arrive[blockIdx.x]).
template <int CLASS, int ZT, bool GIBBS, bool EACC>
__global__ void __launch_bounds__(128)
rows_sweep_kernel(ModelView m, RowsView v, int colour, uint32_t sweep_index, int exact_log, float escale, int want_mags,
                  long long *acc, unsigned int *arrive) {
    extern __shared__ uint32_t lut_sm[];  // FERRO / UNIT: [T][4z + 1] acceptance counts
    constexpr int NS = rows_ns<CLASS>();
    __shared__ unsigned long long blk[NS][2];
    __shared__ int is_last;
    const int z = ZT > 0 ? ZT : m.z, width = 4 * z + 1;
    if (EACC && threadIdx.x < NS * 2) blk[threadIdx.x >> 1][threadIdx.x & 1] = 0ull;
    if (CLASS != COUP_F32) {
        for (int i = threadIdx.x; i < m.T * width; i += blockDim.x) lut_sm[i] = m.lut[i];
    }
    if (EACC || CLASS != COUP_F32) __syncthreads();
    const int sblocks = (m.S + NS - 1) / NS;
    const int64_t d = blockIdx.x / sblocks;
    const int slot0 = (int)(blockIdx.x % sblocks) * NS;
    const int cls = colour % v.m_half;
    const uint32_t n_cls = v.class_start[cls + 1] - v.class_start[cls];
    const uint32_t ci = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = ci < n_cls * (uint32_t)v.kpr;
    if constexpr (!EACC) {
        if (active) rows_sweep_body<CLASS, ZT, GIBBS, false>(m, v, lut_sm, m.spins + d * m.S * m.N, d, slot0, ci, colour, sweep_index, exact_log);
    } else {
    long long e_part[NS];
    int dn_part[NS];
#pragma unroll
    for (int ss = 0; ss < NS; ss++) { e_part[ss] = 0; dn_part[ss] = 0; }
    if (active)
        rows_sweep_body<CLASS, ZT, GIBBS, true>(m, v, lut_sm, m.spins + d * m.S * m.N, d, slot0, ci, colour, sweep_index, exact_log, escale,
                                                e_part, dn_part);
#pragma unroll
    for (int ss = 0; ss < NS; ss++) {  // a 64-bit sum as two REDUX: low 16 bits (unsigned) and the rest (signed)
        const int lo = __reduce_add_sync(0xFFFFFFFFu, (int)(e_part[ss] & 0xFFFF));
        const int hi = __reduce_add_sync(0xFFFFFFFFu, (int)(e_part[ss] >> 16));
        const int dn = __reduce_add_sync(0xFFFFFFFFu, dn_part[ss]);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&blk[ss][0], (unsigned long long)(((long long)hi << 16) + (long long)lo));
            atomicAdd(&blk[ss][1], (unsigned long long)(long long)dn);
        }
    }
    __syncthreads();
    if (threadIdx.x < NS * 2) {
        const int ss = threadIdx.x >> 1, slot = slot0 + ss;
        if (slot < m.S) {
            const int sysl = m.system_ids[d * m.S + slot];
            const int64_t sysg = d * m.S + sysl;
            if (sysl >= m.sys_lo && sysl < m.sys_hi) atomicAdd((unsigned long long *)&acc[2 * sysg + (threadIdx.x & 1)], blk[ss][threadIdx.x & 1]);
        }
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&arrive[blockIdx.x], 1u) == gridDim.y - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < NS) {
        const int slot = slot0 + threadIdx.x;
        const int sysl = slot < m.S ? m.system_ids[d * m.S + slot] : -1;
        if (sysl >= m.sys_lo && sysl < m.sys_hi) {
            const int64_t sysg = d * m.S + sysl;
            const long long e_tot = (long long)atomicExch((unsigned long long *)&acc[2 * sysg], 0ull);
            const long long d_tot = (long long)atomicExch((unsigned long long *)&acc[2 * sysg + 1], 0ull);
            if (CLASS == COUP_F32) m.energies[sysg] = __fdiv_rn((float)((double)e_tot / (double)escale), (float)m.N);
            else m.energies[sysg] = __fdiv_rn((float)e_tot, (float)m.N);
            if (want_mags) m.mags[sysg] = m.N - 2 * d_tot;
        }
    }
    if (threadIdx.x == 0) arrive[blockIdx.x] = 0u;
    }
}

// energies (+ magnetisation sums): grid = (D * S, nb).  Integer classes split a system over nb blocks: partial sums meet
// in acc[2 * sys + {0, 1}] (64-bit integer atomics: order-independent) and the block that arrives last converts them and
// re-zeroes the scratch; fp32 couplings keep one block per system (f64 sums in a fixed order: reproducible).
// e = (sum_i sum_d s_i s_fwd J) / N with the reference's final f32 division (energy.rs:99-108).
template <int CLASS, int ZT>
__global__ void __launch_bounds__(256) rows_energy_kernel(ModelView m, RowsView v, int want_mags, long long *acc, unsigned int *arrive) {
    __shared__ long long sh_ll[32];
    __shared__ double sh_d[32];
    constexpr int ZA = ZT > 0 ? ZT : 16;
    const int64_t sysg = blockIdx.x;
    const int64_t d = sysg / m.S;
    {
        const int sysl = (int)(sysg - d * m.S);
        if (sysl < m.sys_lo || sysl >= m.sys_hi) return;  // system-split handle: another process owns this system
    }
    const int8_t *s = m.spins + sysg * m.N;
    const int z = ZT > 0 ? ZT : m.z, L = v.L;
    const uint32_t n_seg = (uint32_t)(v.n_rows * v.kpr);
    long long isum = 0, dn = 0;  // FERRO: unsatisfied forward bonds; UNIT: sum s s J
    double acc_f = 0.0;
    for (uint32_t ci = blockIdx.y * blockDim.x + threadIdx.x; ci < n_seg; ci += gridDim.y * blockDim.x) {
        uint32_t r;
        int k;
        rows_split(v, ci, r, k);
        const uint64_t C = rows_ld8(s + (int64_t)r * L + 8 * k);
        dn += __popcll(C & 0x8080808080808080ull);
        uint64_t X[ZA];  // sign bit of byte j set where forward bond (site j, direction kk) joins opposite spins
#pragma unroll
        for (int kk = 0; kk < ZA; kk++)
            if (kk < z) X[kk] = C ^ rows_shifted(s + (int64_t)v.nbr_row[((size_t)r * z + kk) * 2] * L, L, k, v.dl[kk]);
        if (CLASS == COUP_FERRO) {
#pragma unroll
            for (int kk = 0; kk < ZA; kk++)
                if (kk < z) isum += __popcll(X[kk] & 0x8080808080808080ull);
        } else {
            const size_t jbase = ((size_t)d * m.N + (size_t)r * L + 8 * k) * z;  // 8 sites x z couplings, contiguous
            if (CLASS == COUP_F32) {
                // 8z consecutive floats (a multiple of 16 bytes from a 16-byte aligned start): 2z vector loads
                float Jv[ZT > 0 ? 8 * ZT : 1];
                const float4 *jp = reinterpret_cast<const float4 *>(m.Jf + jbase);
#pragma unroll
                for (int qv = 0; qv < 2 * ZT; qv++) {
                    {
                        const float4 t4 = __ldg(jp + qv);
                        Jv[4 * qv] = t4.x; Jv[4 * qv + 1] = t4.y; Jv[4 * qv + 2] = t4.z; Jv[4 * qv + 3] = t4.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
#pragma unroll
                    for (int kk = 0; kk < ZA; kk++) {
                        if (kk < z) {
                            const bool opp = ((X[kk] >> (8 * j)) & 0x80u) != 0;
                            const float J = ZT > 0 ? Jv[j * ZT + kk] : m.Jf[jbase + (size_t)j * z + kk];
                            acc_f += (double)(opp ? -J : J);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
#pragma unroll
                    for (int kk = 0; kk < ZA; kk++) {
                        if (kk < z) {
                            const bool opp = ((X[kk] >> (8 * j)) & 0x80u) != 0;
                            const int J = m.J8[jbase + (size_t)j * z + kk];
                            isum += opp ? -J : J;
                        }
                    }
                }
            }
        }
    }
    if (CLASS == COUP_F32) {
        const double tot = block_sum<double>(acc_f, sh_d);
        if (threadIdx.x == 0) m.energies[sysg] = __fdiv_rn((float)tot, (float)m.N);
        if (want_mags) {
            const long long td = block_sum<long long>(dn, sh_ll);
            if (threadIdx.x == 0) m.mags[sysg] = m.N - 2 * td;
        }
        return;
    }
    const long long ti = block_sum<long long>(isum, sh_ll);
    const long long td = block_sum<long long>(dn, sh_ll);
    if (threadIdx.x == 0) {
        long long e_tot = ti, d_tot = td;
        bool last = true;
        if (gridDim.y > 1) {
            atomicAdd((unsigned long long *)&acc[2 * sysg], (unsigned long long)ti);
            atomicAdd((unsigned long long *)&acc[2 * sysg + 1], (unsigned long long)td);
            __threadfence();
            last = atomicAdd(&arrive[sysg], 1u) == gridDim.y - 1;
            if (last) {
                __threadfence();
                e_tot = (long long)atomicExch((unsigned long long *)&acc[2 * sysg], 0ull);
                d_tot = (long long)atomicExch((unsigned long long *)&acc[2 * sysg + 1], 0ull);
                arrive[sysg] = 0u;
            }
        }
        if (last) {
            const long long bonds = CLASS == COUP_FERRO ? (long long)z * m.N - 2 * e_tot : e_tot;
            m.energies[sysg] = __fdiv_rn((float)bonds, (float)m.N);
            if (want_mags) m.mags[sysg] = m.N - 2 * d_tot;
        }
    }
}

// integer overlap dots (overlap.rs:259-281): grid = (D * P * T, nb), same split / last-block scheme as the energies
__global__ void __launch_bounds__(256)
rows_overlap_kernel(ModelView m, RowsView v, long long *dot_spin, long long *dot_link, long long *acc, unsigned int *arrive) {
    __shared__ long long sh[32];
    const int64_t idx = blockIdx.x;  // (d*P + p)*T + t
    const int t = (int)(idx % m.T);
    const int p = (int)((idx / m.T) % m.P);
    const int64_t d = idx / ((int64_t)m.T * m.P);
    const int sa = m.system_ids[d * m.S + (2 * p) * m.T + t];
    const int sb = m.system_ids[d * m.S + (2 * p + 1) * m.T + t];
    const int8_t *a = m.spins + (d * m.S + sa) * m.N;
    const int8_t *b = m.spins + (d * m.S + sb) * m.N;
    const int z = m.z, L = v.L;
    const uint32_t n_seg = (uint32_t)(v.n_rows * v.kpr);
    long long neg_q = 0, neg_l = 0;  // sites with q_i = -1, links with q_i q_j = -1
    for (uint32_t ci = blockIdx.y * blockDim.x + threadIdx.x; ci < n_seg; ci += gridDim.y * blockDim.x) {
        uint32_t r;
        int k;
        rows_split(v, ci, r, k);
        const int64_t o = (int64_t)r * L + 8 * k;
        const uint64_t x = rows_ld8(a + o) ^ rows_ld8(b + o);  // sign bit set where the replicas differ
        neg_q += __popcll(x & 0x8080808080808080ull);
        for (int kk = 0; kk < z; kk++) {
            const int64_t nro = (int64_t)v.nbr_row[((size_t)r * z + kk) * 2] * L;
            const uint64_t xf = rows_shifted(a + nro, L, k, v.dl[kk]) ^ rows_shifted(b + nro, L, k, v.dl[kk]);
            neg_l += __popcll((x ^ xf) & 0x8080808080808080ull);
        }
    }
    const long long tq = block_sum<long long>(neg_q, sh);
    const long long tl = block_sum<long long>(neg_l, sh);
    if (threadIdx.x == 0) {
        long long q_tot = tq, l_tot = tl;
        bool last = true;
        if (gridDim.y > 1) {
            atomicAdd((unsigned long long *)&acc[2 * idx], (unsigned long long)tq);
            atomicAdd((unsigned long long *)&acc[2 * idx + 1], (unsigned long long)tl);
            __threadfence();
            last = atomicAdd(&arrive[idx], 1u) == gridDim.y - 1;
            if (last) {
                __threadfence();
                q_tot = (long long)atomicExch((unsigned long long *)&acc[2 * idx], 0ull);
                l_tot = (long long)atomicExch((unsigned long long *)&acc[2 * idx + 1], 0ull);
                arrive[idx] = 0u;
            }
        }
        if (last) {
            dot_spin[idx] = m.N - 2 * q_tot;
            dot_link[idx] = (long long)z * m.N - 2 * l_tot;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Resident kernel for small realizations (README quickstart sizes: 32 systems x 1024 spins): ONE launch runs many sweeps.
// One CTA owns realization d: its [S][N] spins live in shared memory for the whole launch, and the per-sweep sequence of
// simulation/mod.rs:405-432, 486-509, 527-529, 748-796 -- colour passes, energies (+ magnetisations), overlap dots, fold,
// parallel tempering -- runs inside the CTA with __syncthreads() where the multi-launch path has kernel boundaries
// (5.5 launches per sweep made those sizes launch-bound).  Integer coupling classes only (their sums do not depend on the
// order of addition, so every result is bit-identical to the multi-launch path).
struct ResidentArgs {
    int64_t sweep_id0;       // index of the first sweep inside this sample() call
    int n_sweeps;
    int64_t warmup_sweeps;   // record = sweep_id >= warmup_sweeps (mod.rs:410)
    int64_t pt_interval;     // 0: no parallel tempering
    int pt_schedule;
    uint32_t sweep_counter0; // RNG-SPEC sweep index of the first sweep
    uint32_t pt_event0;
    int parity0;
    int spins_in_smem;
    long long *dot_spin, *dot_link;  // [D][P][T] (the StatsView holds the same arrays read-only)
};

constexpr int RESIDENT_THREADS = 512;

// bytes of the per-realization scalar state the resident kernel keeps in shared memory (8-byte aligned block)
__host__ __device__ inline size_t resident_scalar_bytes(int S, int T, int P) {
    const size_t n_edges = T > 1 ? (size_t)(T - 1) : 1;
    size_t b = 8 * ((size_t)11 * T + (size_t)S + 2 * (size_t)P * T + 2 * n_edges + (size_t)S);  // sums, mags, dots, edge counters, round trips
    b += 4 * ((size_t)S + (size_t)T + (size_t)S);                                              // energies, temps, system ids
    b += (size_t)S;                                                                            // trip states
    return (b + 15) & ~size_t(15);
}

template <int CLASS, int ZT, bool GIBBS>
__global__ void __launch_bounds__(RESIDENT_THREADS)
rows_resident_kernel(ModelView mg, RowsView vg, StatsView stg, PtView ptg, ResidentArgs a) {
    extern __shared__ __align__(16) uint32_t res_sm[];
    const int z = ZT > 0 ? ZT : mg.z, width = 4 * z + 1, L = vg.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = RESIDENT_THREADS / 32;
    const int64_t dg = blockIdx.x;
    const int S = mg.S, T = mg.T, P = mg.P, n_edges = T > 1 ? T - 1 : 1;
    uint32_t *lut_sm = res_sm;
    // ---- the realization's scalar state (system ids, energies, magnetisation sums, running sums, pair dots, exchange counters)
    // lives in shared memory for the whole launch, so no phase of a sweep waits on an L2 round trip; m / v / st / pt below are
    // views of that copy with the realization index folded in (d = 0); the copy goes back to global memory at the end of the
    // launch.  (Measured at the README quickstart: 98 ms per 5000 sweeps before and after — the sweep is bound by the chain of
    // short phases on one SM, not by those round trips; kept because it takes the global arrays off the per-sweep path.)
    unsigned char *sc = reinterpret_cast<unsigned char *>(res_sm + ((T * width + 3) & ~3));
    double *sums_sm = reinterpret_cast<double *>(sc);
    long long *mag_sm = reinterpret_cast<long long *>(sums_sm + 11 * T);
    long long *dsp_sm = mag_sm + S, *dlk_sm = dsp_sm + P * T;
    unsigned long long *ea_sm = reinterpret_cast<unsigned long long *>(dlk_sm + P * T), *eacc_sm = ea_sm + n_edges, *rt_sm = eacc_sm + n_edges;
    float *en_sm = reinterpret_cast<float *>(rt_sm + S), *temps_sm = en_sm + S;
    int32_t *sid_sm = reinterpret_cast<int32_t *>(temps_sm + T);
    uint8_t *trip_sm = reinterpret_cast<uint8_t *>(sid_sm + S);
    int8_t *g_spins = mg.spins + dg * S * mg.N;
    const int64_t n_bytes = (int64_t)S * mg.N;
    int8_t *spins_d = a.spins_in_smem ? reinterpret_cast<int8_t *>(sc + resident_scalar_bytes(S, T, P)) : g_spins;
    const int64_t bins = mg.N + 1;
    for (int i = tid; i < T * width; i += RESIDENT_THREADS) lut_sm[i] = mg.lut[i];
    for (int i = tid; i < 11 * T; i += RESIDENT_THREADS) sums_sm[i] = stg.sums[dg * 11 * T + i];
    for (int i = tid; i < S; i += RESIDENT_THREADS) {
        mag_sm[i] = mg.mags[dg * S + i];
        en_sm[i] = mg.energies[dg * S + i];
        sid_sm[i] = mg.system_ids[dg * S + i];
        rt_sm[i] = ptg.round_trips[dg * S + i];
        trip_sm[i] = ptg.trip_state[dg * S + i];
    }
    for (int i = tid; i < P * T; i += RESIDENT_THREADS) {
        dsp_sm[i] = a.dot_spin[dg * P * T + i];
        dlk_sm[i] = a.dot_link[dg * P * T + i];
    }
    for (int i = tid; i < T - 1; i += RESIDENT_THREADS) {
        ea_sm[i] = ptg.edge_attempts[dg * (T - 1) + i];
        eacc_sm[i] = ptg.edge_acceptances[dg * (T - 1) + i];
    }
    for (int i = tid; i < T; i += RESIDENT_THREADS) temps_sm[i] = mg.temps[i];
    if (a.spins_in_smem)
        for (int64_t i = tid; i < n_bytes / 16; i += RESIDENT_THREADS)
            reinterpret_cast<uint4 *>(spins_d)[i] = reinterpret_cast<const uint4 *>(g_spins)[i];
    ModelView m = mg;
    m.D = 1;
    m.sample_offset = mg.sample_offset + dg;  // seeds stay those of the global realization index
    m.system_ids = sid_sm;
    m.energies = en_sm;
    m.mags = mag_sm;
    m.temps = temps_sm;
    m.J8 = mg.J8 ? mg.J8 + (size_t)dg * mg.N * z : nullptr;
    RowsView v = vg;
    v.keys = vg.keys + dg;
    StatsView st = stg;
    st.sums = sums_sm;
    if (stg.hist) {
        st.hist = stg.hist + dg * T * bins;
        st.ql_at_q = stg.ql_at_q + dg * T * bins;
        st.ql2_at_q = stg.ql2_at_q + dg * T * bins;
    }
    st.dot_spin = dsp_sm;
    st.dot_link = dlk_sm;
    PtView pt = ptg;
    pt.edge_attempts = ea_sm;
    pt.edge_acceptances = eacc_sm;
    pt.round_trips = rt_sm;
    pt.trip_state = trip_sm;
    a.dot_spin = dsp_sm;
    a.dot_link = dlk_sm;
    const int64_t d = 0;
    __syncthreads();
    constexpr int NS = rows_ns<CLASS>();
    const int sblocks = (m.S + NS - 1) / NS;
    const uint32_t n_seg = (uint32_t)(v.n_rows * v.kpr);
    uint32_t pt_event = a.pt_event0;
    int parity = a.parity0;
    for (int sw = 0; sw < a.n_sweeps; sw++) {
        const int64_t sid = a.sweep_id0 + sw;
        const uint32_t sweep_index = a.sweep_counter0 + (uint32_t)sw;
        for (int col = 0; col < m.n_colours; col++) {
            const int cls = col % v.m_half;
            const uint32_t nseg = (v.class_start[cls + 1] - v.class_start[cls]) * (uint32_t)v.kpr;
            for (uint32_t it = tid; it < (uint32_t)sblocks * nseg; it += RESIDENT_THREADS)
                rows_sweep_body<CLASS, ZT, GIBBS>(m, v, lut_sm, spins_d, d, (int)(it / nseg) * NS, it % nseg, col, sweep_index, 0);
            __syncthreads();
        }
        const bool record = sid >= a.warmup_sweeps;
        const bool pt_this = a.pt_interval > 0 && sid % a.pt_interval == 0;
        if (record || pt_this) {  // mod.rs:486-509; energy.rs:99-108 per system, one warp each
            for (int sys = warp; sys < m.S; sys += n_warps) {
                const int8_t *s = spins_d + (int64_t)sys * m.N;
                long long isum = 0, dn = 0;  // FERRO: unsatisfied forward bonds; UNIT: sum s s J
                for (uint32_t ci = lane; ci < n_seg; ci += 32) {
                    uint32_t r;
                    int k;
                    rows_split(v, ci, r, k);
                    const uint64_t C = rows_ld8(s + (int64_t)r * L + 8 * k);
                    dn += __popcll(C & 0x8080808080808080ull);
                    for (int kk = 0; kk < z; kk++) {
                        const uint64_t X = C ^ rows_shifted(s + (int64_t)v.nbr_row[((size_t)r * z + kk) * 2] * L, L, k, v.dl[kk]);
                        if (CLASS == COUP_FERRO) {
                            isum += __popcll(X & 0x8080808080808080ull);
                        } else {
                            const size_t jbase = ((size_t)d * m.N + (size_t)r * L + 8 * k) * z;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int J = m.J8[jbase + (size_t)j * z + kk];
                                isum += ((X >> (8 * j)) & 0x80u) ? -J : J;
                            }
                        }
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    isum += __shfl_xor_sync(0xFFFFFFFFu, isum, o);
                    dn += __shfl_xor_sync(0xFFFFFFFFu, dn, o);
                }
                if (lane == 0) {
                    const long long bonds = CLASS == COUP_FERRO ? (long long)z * m.N - 2 * isum : isum;
                    m.energies[d * m.S + sys] = __fdiv_rn((float)bonds, (float)m.N);
                    if (record) m.mags[d * m.S + sys] = m.N - 2 * dn;
                }
            }
            __syncthreads();
        }
        if (record) {
            if (m.P > 0) {  // overlap.rs:259-281 with this sweep's pre-exchange system_ids, one warp per (pair, slot)
                for (int idx = warp; idx < m.P * m.T; idx += n_warps) {
                    const int t = idx % m.T, p = idx / m.T;
                    const int8_t *sa = spins_d + (int64_t)m.system_ids[d * m.S + (2 * p) * m.T + t] * m.N;
                    const int8_t *sb = spins_d + (int64_t)m.system_ids[d * m.S + (2 * p + 1) * m.T + t] * m.N;
                    long long neg_q = 0, neg_l = 0;
                    for (uint32_t ci = lane; ci < n_seg; ci += 32) {
                        uint32_t r;
                        int k;
                        rows_split(v, ci, r, k);
                        const int64_t o = (int64_t)r * L + 8 * k;
                        const uint64_t x = rows_ld8(sa + o) ^ rows_ld8(sb + o);
                        neg_q += __popcll(x & 0x8080808080808080ull);
                        for (int kk = 0; kk < z; kk++) {
                            const int64_t nro = (int64_t)v.nbr_row[((size_t)r * z + kk) * 2] * L;
                            const uint64_t xf = rows_shifted(sa + nro, L, k, v.dl[kk]) ^ rows_shifted(sb + nro, L, k, v.dl[kk]);
                            neg_l += __popcll((x ^ xf) & 0x8080808080808080ull);
                        }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        neg_q += __shfl_xor_sync(0xFFFFFFFFu, neg_q, o);
                        neg_l += __shfl_xor_sync(0xFFFFFFFFu, neg_l, o);
                    }
                    if (lane == 0) {
                        a.dot_spin[(d * m.P + p) * m.T + t] = m.N - 2 * neg_q;
                        a.dot_link[(d * m.P + p) * m.T + t] = (long long)z * m.N - 2 * neg_l;
                    }
                }
                __syncthreads();
            }
            for (int t = tid; t < m.T; t += RESIDENT_THREADS)  // mod.rs:543-578
                fold_one<0, true>(
                    m, st, d, t, m.P > 0,
                    [&](int r) { return m.mags[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
                    [&](int r) { return m.energies[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
                    [&](int p) { return a.dot_spin[(d * m.P + p) * m.T + t]; },
                    [&](int p) { return a.dot_link[(d * m.P + p) * m.T + t]; });
            __syncthreads();
        }
        if (pt_this) {  // mod.rs:748-796
            if (m.T >= 2) {
                if (tid < m.R) pt_exchange_body(m, pt, d, tid, a.pt_schedule, parity, pt_event);
                if (a.pt_schedule == 1) parity = 1 - parity;
                __syncthreads();
            }
            pt_event++;
        }
    }
    __syncthreads();
    for (int i = tid; i < 11 * T; i += RESIDENT_THREADS) stg.sums[dg * 11 * T + i] = sums_sm[i];
    for (int i = tid; i < S; i += RESIDENT_THREADS) {
        mg.mags[dg * S + i] = mag_sm[i];
        mg.energies[dg * S + i] = en_sm[i];
        mg.system_ids[dg * S + i] = sid_sm[i];
        ptg.round_trips[dg * S + i] = rt_sm[i];
        ptg.trip_state[dg * S + i] = trip_sm[i];
    }
    for (int i = tid; i < P * T; i += RESIDENT_THREADS) {
        // (a.dot_spin / a.dot_link were redirected to the shared copy above: the handle's arrays are the StatsView's)
        const_cast<long long *>(stg.dot_spin)[dg * P * T + i] = dsp_sm[i];
        const_cast<long long *>(stg.dot_link)[dg * P * T + i] = dlk_sm[i];
    }
    for (int i = tid; i < T - 1; i += RESIDENT_THREADS) {
        ptg.edge_attempts[dg * (T - 1) + i] = ea_sm[i];
        ptg.edge_acceptances[dg * (T - 1) + i] = eacc_sm[i];
    }
    if (a.spins_in_smem)
        for (int64_t i = tid; i < n_bytes / 16; i += RESIDENT_THREADS)
            reinterpret_cast<uint4 *>(g_spins)[i] = reinterpret_cast<const uint4 *>(spins_d)[i];
}
#endif  // __CUDACC__

}  // namespace pp
