// pp_kernels_prows.cuh — ferromagnets on any row-alternating lattice with ONE BIT per spin, a whole system resident in shared memory.
//
// The generalisation of the bit-packed slab kernel (pp_kernels_slabp.cuh) to the lattices the row tables of pp_kernels_rows.cuh
// describe: any neighbour offsets whose last component is -1, 0 or +1 and whose colouring c(x) = (sum_d a_d x_d) mod m has
// a_last = m / 2 (square, cubic, triangular, FCC, BCC ...), last extent a multiple of 64.  BASELINE configs[2] (128 triangular
// systems of 256 x 256 sites, heat-bath sweeps) is the case it was built for: with one byte per spin a system is 64 KiB, a sweep is
// seven launch-latency-sized launches over L2-resident data; with one bit per spin a system is 8 KiB, ONE CTA keeps it in shared
// memory for all colour classes of every sweep of a launch, and the per-sweep launch sequence shrinks to sweep (+ energies) /
// overlap / fold.
//
// Replaces, for ferromagnets on this layout,
//   metropolis_sweep / gibbs_sweep (lookup rule)     spin-sim/src/mcmc/sweep.rs:170-185, 220-284
//   compute_energies_and_magnetizations_into         spin-sim/src/spins/energy.rs:59-110
//   OverlapAccum::collect (integer dots)             spin-sim/src/statistics/overlap.rs:259-281
//
// Storage of one system: u32 [rows][2][W], W = L / 64: row r (a fixed value of the leading coordinates) keeps its even-x and its
// odd-x sites as two sets of W words, bit b of word w of set p = site x = 64 w + 2 b + p, bit = 1: spin -1.  In colour pass c a
// row of class (c mod m/2) updates one of its two sets — the even one if its phase row_a equals c, else the odd one.  The
// neighbour in direction k of the 32 sites of a word is one word of row nbr_row[r][k]: the same set at the same position when the
// offset's last component is 0, the other set at the same position or shifted by one bit (completed by the adjacent word) when it
// is +-1.  Bond words are XORs (ferromagnet), counted bit-sliced; the acceptance rule  flip <=> draw < table[t][2 unsat]
// (sweep.rs:182-184) is evaluated for 32 sites at once from per-site threshold masks, mask_u bit b = [draw_b < table[t][2u]]:
//   flip = OR_u ( [unsat == u] & mask_u ).
// Draws: RNG-SPEC v2 packed mapping (pp_rng.cuh) — the 32 sites of a word are the ranks 32 q .. 32 q + 31 of their colour class,
// q = row_ord[r] * W + w, and share six Philox calls.
#pragma once
#include "pp_device.cuh"
#include "pp_kernels_rows.cuh"
#include "pp_kernels_stats.cuh"

namespace pp {

struct PRowsView {
    uint32_t *words;   // [D][S][rows][2][W] by system
    int W;             // words per row and parity set
    int64_t sys_words; // rows * 2 * W
};

constexpr int PROWS_THREADS = 512;
constexpr int PROWS_MAX_Z = 4;  // forward directions (z2 = 2z <= 8 bond words)

// word of (row nr, the set / shift the neighbour in a direction with last component `dls` of the sites of set `p`, word w) needs
__device__ __forceinline__ uint32_t prows_nbr_word(const uint32_t *sys, const int W, const uint32_t nr, const int p, const int w, const int dls) {
    const uint32_t *row = sys + (size_t)nr * 2 * W;
    if (dls == 0) return row[p * W + w];
    const uint32_t *oth = row + (1 - p) * W;
    const uint32_t O = oth[w];
    if (dls > 0) {  // x + 1: an even site's neighbour is the odd site of the same pair; an odd site's is the next pair's even site
        if (p == 0) return O;
        return (O >> 1) | (oth[w + 1 == W ? 0 : w + 1] << 31);
    }
    if (p == 1) return O;  // x - 1: an odd site's neighbour is the even site of the same pair
    return (O << 1) | (oth[w ? w - 1 : W - 1] >> 31);
}

// bit-sliced count of up to 8 bond words: u[0..3] = bits of the per-lane number of set words
template <int Z2>
__device__ __forceinline__ void prows_count(const uint32_t *b, uint32_t (&u)[4]) {
    u[0] = u[1] = u[2] = u[3] = 0u;
#pragma unroll
    for (int k = 0; k < Z2; k++) {
        uint32_t c = b[k], t;
        t = u[0] & c; u[0] ^= c; c = t;
        t = u[1] & c; u[1] ^= c; c = t;
        t = u[2] & c; u[2] ^= c; c = t;
        u[3] ^= c;
    }
}

// grid = D * S CTAs: CTA (d, slot) keeps the words of system system_ids[d][slot] in shared memory for `n_sweeps` sweeps (every
// colour class of each) and, when want_energy, leaves that system's energy (+ magnetisation sum) behind; n_sweeps = 0 with
// want_energy is the plain energy evaluation.  Z = forward directions, NM = thresholds compared per site: Z (Metropolis: the
// counts for unsat >= Z are 2^24 and the others are below 2^24, host-checked) or 2 Z + 1 (any table).
// dynamic shared memory: sys_words words.
template <int Z, int NM>
__global__ void __launch_bounds__(PROWS_THREADS)
prows_sweep_kernel(ModelView m, RowsView v, PRowsView pv, uint32_t sweep_index, int n_sweeps, int want_energy, int want_mags) {
    extern __shared__ __align__(16) uint32_t prows_sm[];
    __shared__ uint32_t thr_sm[2 * PROWS_MAX_Z + 1];
    __shared__ long long red_sm[32];
    constexpr int Z2 = 2 * Z;
    const int tid = threadIdx.x;
    const int64_t d = blockIdx.x / m.S;
    const int slot = (int)(blockIdx.x % m.S);
    const int t = slot % m.T;  // realization.rs:166
    const int sysl = m.system_ids[d * m.S + slot];  // parallel.rs:27-33: spins by system, temperature by slot
    if (sysl < m.sys_lo || sysl >= m.sys_hi) return;  // system-split handle: another process updates this system
    const int64_t sysg = d * m.S + sysl;
    uint32_t *gw = pv.words + sysg * pv.sys_words;
    const int W = pv.W;
    for (int64_t i = tid; i < pv.sys_words / 4; i += PROWS_THREADS) reinterpret_cast<uint4 *>(prows_sm)[i] = reinterpret_cast<const uint4 *>(gw)[i];
    if (tid <= Z2) thr_sm[tid] = m.lut[t * (4 * Z + 1) + 2 * tid];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
    __syncthreads();
    uint32_t T[NM];
#pragma unroll
    for (int u = 0; u < NM; u++) T[u] = NM == Z ? thr_sm[u] << 8 : thr_sm[u];
    const uint64_t key = v.keys[d];
    const PhiloxKeys ks = philox_keys((uint32_t)key, (uint32_t)(key >> 32));
    int dls[Z];
#pragma unroll
    for (int k = 0; k < Z; k++) dls[k] = v.dl[k];

    for (int sw = 0; sw < n_sweeps; sw++) {
        for (int colour = 0; colour < m.n_colours; colour++) {
            const int cls = colour % v.m_half;
            const uint32_t row0 = v.class_start[cls], n_items = (v.class_start[cls + 1] - row0) * (uint32_t)W;
            const uint32_t tagc = TAG_SWEEP_PACKED | (uint32_t)colour;
            for (uint32_t it = tid; it < n_items; it += PROWS_THREADS) {
                const uint32_t ri = it / (uint32_t)W;
                const int w = (int)(it - ri * (uint32_t)W);
                const uint32_t r = v.class_rows[row0 + ri];
                const int p = (int)v.row_a[r] == colour ? 0 : 1;  // the set of this row that has colour `colour`
                uint32_t *slf = prows_sm + ((size_t)r * 2 + p) * W + w;
                const uint32_t C = *slf;
                uint32_t b[Z2];
#pragma unroll
                for (int k = 0; k < Z; k++) {
                    b[2 * k] = C ^ prows_nbr_word(prows_sm, W, v.nbr_row[((size_t)r * Z + k) * 2], p, w, dls[k]);
                    b[2 * k + 1] = C ^ prows_nbr_word(prows_sm, W, v.nbr_row[((size_t)r * Z + k) * 2 + 1], p, w, -dls[k]);
                }
                uint32_t un[4];
                prows_count<Z2>(b, un);
                uint32_t M[NM];
#pragma unroll
                for (int u = 0; u < NM; u++) M[u] = 0u;
                const uint32_t q = v.row_ord[r] * (uint32_t)W + (uint32_t)w;  // rank >> 5 of the word's sites
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t wd[12];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const u32x4 o = philox4x32_k(q, sweep_index + (uint32_t)sw, (uint32_t)sysl, tagc | ((uint32_t)(3 * h + c) << 8), ks);
                        wd[4 * c] = o.x; wd[4 * c + 1] = o.y; wd[4 * c + 2] = o.z; wd[4 * c + 3] = o.w;
                    }
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const uint32_t A = wd[3 * g], B = wd[3 * g + 1], Cw = wd[3 * g + 2];
                        const uint32_t y3 = __byte_perm(__byte_perm(Cw, B, 0x0400), A, 0x4210);
                        const uint32_t ys[4] = {A, B, Cw, y3};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t bit = 1u << (16 * h + 4 * g + j);
                            const uint32_t y = NM == Z ? ys[j] : ys[j] >> 8;
#pragma unroll
                            for (int u = 0; u < NM; u++)
                                if (y < T[u]) M[u] |= bit;
                        }
                    }
                }
                // flip = OR_u ([unsat == u] & M_u); unsat >= NM always flips when NM == Z (Metropolis, energy change <= 0)
                uint32_t flip = 0u;
#pragma unroll
                for (int u = 0; u <= Z2; u++) {
                    const uint32_t eq = ((u & 1) ? un[0] : ~un[0]) & ((u & 2) ? un[1] : ~un[1]) & ((u & 4) ? un[2] : ~un[2]) & ((u & 8) ? un[3] : ~un[3]);
                    flip |= u < NM ? (eq & M[u]) : eq;
                }
                *slf = C ^ flip;
            }
            __syncthreads();
        }
    }
    if (want_energy) {  // energy.rs:99-108: every bond once through its forward direction; down spins of both sets
        long long unsat = 0, dn = 0;
        const uint32_t n_words = (uint32_t)pv.sys_words;
        for (uint32_t i = tid; i < n_words; i += PROWS_THREADS) {
            const uint32_t r = i / (uint32_t)(2 * W);
            const int p = (int)((i / (uint32_t)W) & 1u), w = (int)(i % (uint32_t)W);
            const uint32_t C = prows_sm[i];
            dn += __popc(C);
#pragma unroll
            for (int k = 0; k < Z; k++) unsat += __popc(C ^ prows_nbr_word(prows_sm, W, v.nbr_row[((size_t)r * Z + k) * 2], p, w, dls[k]));
        }
        const long long tu = block_sum<long long>(unsat, red_sm);
        const long long td = block_sum<long long>(dn, red_sm);
        if (tid == 0) {
            m.energies[sysg] = __fdiv_rn((float)((long long)Z * m.N - 2 * tu), (float)m.N);
            if (want_mags) m.mags[sysg] = m.N - 2 * td;
        }
    }
    if (n_sweeps > 0)
        for (int64_t i = tid; i < pv.sys_words / 4; i += PROWS_THREADS) reinterpret_cast<uint4 *>(gw)[i] = reinterpret_cast<const uint4 *>(prows_sm)[i];
}

// integer overlap dots (overlap.rs:259-281) of the replica pairs at one (realization, temperature), words straight from global
// memory / L2, and — want_fold — the recorded-sweep fold of that (d, t) right behind them (simulation/mod.rs:543-578,
// statistics/overlap.rs:283-306) by the CTA's first thread: one launch instead of two per recorded sweep.  grid = D * T CTAs.
template <int Z>
__global__ void __launch_bounds__(PROWS_THREADS)
prows_overlap_kernel(ModelView m, RowsView v, PRowsView pv, StatsView st, long long *dot_spin, long long *dot_link, int want_fold) {
    __shared__ long long red_sm[32];
    const int64_t d = blockIdx.x / m.T;
    const int t = (int)(blockIdx.x % m.T);
    const int W = pv.W;
    int dls[Z];
#pragma unroll
    for (int k = 0; k < Z; k++) dls[k] = v.dl[k];
    for (int pr = 0; pr < m.P; pr++) {
        const int sa = m.system_ids[d * m.S + (2 * pr) * m.T + t];
        const int sb = m.system_ids[d * m.S + (2 * pr + 1) * m.T + t];
        const uint32_t *a = pv.words + (d * m.S + sa) * pv.sys_words;
        const uint32_t *b = pv.words + (d * m.S + sb) * pv.sys_words;
        long long neg_q = 0, neg_l = 0;
        for (uint32_t i = threadIdx.x; i < (uint32_t)pv.sys_words; i += PROWS_THREADS) {
            const uint32_t r = i / (uint32_t)(2 * W);
            const int p = (int)((i / (uint32_t)W) & 1u), w = (int)(i % (uint32_t)W);
            const uint32_t x = a[i] ^ b[i];  // bit set where the replicas differ
            neg_q += __popc(x);
#pragma unroll
            for (int k = 0; k < Z; k++) {
                const uint32_t nr = v.nbr_row[((size_t)r * Z + k) * 2];
                neg_l += __popc(x ^ prows_nbr_word(a, W, nr, p, w, dls[k]) ^ prows_nbr_word(b, W, nr, p, w, dls[k]));
            }
        }
        const long long tq = block_sum<long long>(neg_q, red_sm);
        const long long tl = block_sum<long long>(neg_l, red_sm);
        if (threadIdx.x == 0) {
            const int64_t idx = (d * m.P + pr) * m.T + t;
            dot_spin[idx] = m.N - 2 * tq;
            dot_link[idx] = (long long)Z * m.N - 2 * tl;
        }
    }
    if (want_fold && threadIdx.x == 0)  // the dots above were written by this thread: program order makes them visible
        fold_one<0, true>(
            m, st, d, t, m.P > 0,
            [&](int r) { return m.mags[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
            [&](int r) { return m.energies[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
            [&](int p) { return dot_spin[(d * m.P + p) * m.T + t]; },
            [&](int p) { return dot_link[(d * m.P + p) * m.T + t]; });
}

// int8 [D][S][N] (+-1, system-major: the layout every other kernel and the API use) <-> packed words; dir 0: pack, 1: unpack.
// One thread = 64 consecutive sites of a row; grid = (D * S, ceil(rows * W / PROWS_THREADS)).
__global__ void __launch_bounds__(PROWS_THREADS) prows_convert_kernel(ModelView m, RowsView v, PRowsView pv, int dir) {
    const int64_t i = (int64_t)blockIdx.y * PROWS_THREADS + threadIdx.x;
    if (i >= v.n_rows * pv.W) return;
    const int64_t sysg = blockIdx.x;
    const int64_t r = i / pv.W;
    const int w = (int)(i % pv.W);
    int8_t *s = m.spins + sysg * m.N + r * v.L + 64 * w;
    uint32_t *row = pv.words + sysg * pv.sys_words + (size_t)r * 2 * pv.W;
    if (dir == 0) {
        uint32_t even = 0u, odd = 0u;
        for (int b = 0; b < 32; b++) {
            if (s[2 * b] < 0) even |= 1u << b;
            if (s[2 * b + 1] < 0) odd |= 1u << b;
        }
        row[w] = even;
        row[pv.W + w] = odd;
    } else {
        const uint32_t even = row[w], odd = row[pv.W + w];
        for (int b = 0; b < 32; b++) {
            s[2 * b] = (even >> b) & 1u ? (int8_t)-1 : (int8_t)1;
            s[2 * b + 1] = (odd >> b) & 1u ? (int8_t)-1 : (int8_t)1;
        }
    }
}

}  // namespace pp
