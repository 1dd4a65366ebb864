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
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#endif

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
constexpr int PROWS_SWEEP_THREADS = 1024;  // sweep kernel: a word's two 16-site halves go to two threads (8 warps per scheduler)
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
// BS (Z = 3, NM = 7; host-checked: the seven counts grow with unsat and stay below 2^24): the number L of thresholds above a site's
// draw comes from a three-step binary search, and  draw < table[t][2 unsat]  <=>  unsat + L >= 7  is a three-LOP3 carry chain per word.
// A work item is HALF a word (16 sites = half a block of the packed draw mapping = three Philox calls): 1024 threads per CTA keep
// eight warps per scheduler busy where one word per thread left four (C3: one item per thread and colour class).
// CLUSTER MODE (cl.on; launched as thread-block clusters of R CTAs = the R replicas sitting at one temperature slot, CTA rank =
// replica): the sweeps sw >= cl.rec_from of the launch are RECORDED inside the kernel — energies and magnetisation sums per system,
// the pair dots of replicas (2p, 2p + 1) by the even-ranked CTA, which reads its partner's words through distributed shared memory,
// and the fold of simulation/mod.rs:543-578 / statistics/overlap.rs:283-306 by the first thread of rank 0, which collects the R
// systems' scalars through distributed shared memory — with two cluster barriers per recorded sweep where the multi-launch path has
// two kernel boundaries and a round trip of the words through global memory.  A whole sample() call without exchanges is ONE launch.
// dynamic shared memory: sys_words words (cluster mode: twice that, the second half holds the pair's XOR words).
struct PRowsCluster {
    int on;             // 0: plain launch (grid = D * S, CTA = slot); 1: cluster launch (grid = D * T * R, CTA = (d, t, r))
    int rec_from;       // sweeps sw >= rec_from of this launch are recorded
    int stage_tables;   // the row tables (nbr_row, row_ord, class_rows, row_a) are copied into shared memory behind the words: a cluster
                        // barrier invalidates L1, and every work item starts with two dependent table reads
    StatsView st;
    long long *dot_spin, *dot_link;
};

#ifdef PP_PROWS_TIMING  // phase clocks of CTA 0 (tools/c3_phases.py; never defined in the product build)
__device__ unsigned long long pp_prows_clk[8];
#define PROWS_CLK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const unsigned long long now_ = clock64(); pp_prows_clk[k] += now_ - clk_last_; clk_last_ = now_; } } while (0)
#else
#define PROWS_CLK(k) do { } while (0)
#endif

template <int Z, int NM, bool BS = false>
__global__ void __launch_bounds__(PROWS_SWEEP_THREADS)
prows_sweep_kernel(ModelView m, RowsView vg, PRowsView pv, uint32_t sweep_index, int n_sweeps, int want_energy, int want_mags, PRowsCluster cl) {
    extern __shared__ __align__(16) uint32_t prows_sm[];
    __shared__ uint32_t thr_sm[2 * PROWS_MAX_Z + 1];
    __shared__ unsigned int cnt_sm[4];                // integer partial sums of the reductions (zero between uses; below 2^31:
                                                      // a system that fits shared memory has fewer than 2^22 sites)
    __shared__ double sums_sm[11];                    // cluster mode, rank 0: running sums of (d, t) for the whole launch
    // cluster mode, rank 0's copy: the R systems' energies and magnetisation sums and the P pairs' dots of up to FOLD_K recorded
    // sweeps (two batches: one being filled, one being folded), stored into it by the owning CTAs through distributed shared memory
    __shared__ float all_e[2][FOLD_K][8];
    __shared__ long long all_m[2][FOLD_K][8], all_q[2][FOLD_K][4], all_l[2][FOLD_K][4];
    __shared__ double fold_tab[FOLD_K][8][11];
    RowsView v = vg;
    constexpr int Z2 = 2 * Z;
    static_assert(!BS || (Z == 3 && NM == 7), "the binary search is written for seven thresholds");
    constexpr bool SHIFTED = NM == Z || BS;  // every compared count is below 2^24: compare the 32-bit words against count << 8
    const int tid = threadIdx.x;
    const int rank = cl.on ? (int)(blockIdx.x % m.R) : 0;  // cluster mode: the replica
    const int64_t d = cl.on ? blockIdx.x / ((int64_t)m.R * m.T) : blockIdx.x / m.S;
    const int slot = cl.on ? rank * m.T + (int)((blockIdx.x / m.R) % m.T) : (int)(blockIdx.x % m.S);
    const int t = slot % m.T;  // realization.rs:166
    const int sysl = m.system_ids[d * m.S + slot];  // parallel.rs:27-33: spins by system, temperature by slot
    if (!cl.on && (sysl < m.sys_lo || sysl >= m.sys_hi)) return;  // system-split handle: another process updates this system
    const int64_t sysg = d * m.S + sysl;
    uint32_t *gw = pv.words + sysg * pv.sys_words;
    const int W = pv.W;
    for (int64_t i = tid; i < pv.sys_words / 4; i += PROWS_SWEEP_THREADS) reinterpret_cast<uint4 *>(prows_sm)[i] = reinterpret_cast<const uint4 *>(gw)[i];
    if (tid <= Z2) thr_sm[tid] = m.lut[t * (4 * Z + 1) + 2 * tid];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
    if (tid < 4) cnt_sm[tid] = 0u;
    if (cl.stage_tables) {
        const uint32_t nr = (uint32_t)vg.n_rows;
        uint32_t *s_nbr = prows_sm + (size_t)pv.sys_words * (cl.on ? 2 : 1), *s_ord = s_nbr + (size_t)nr * Z2, *s_crow = s_ord + nr;
        uint8_t *s_a = reinterpret_cast<uint8_t *>(s_crow + nr);
        for (uint32_t i = tid; i < nr * (uint32_t)Z2; i += PROWS_SWEEP_THREADS) s_nbr[i] = vg.nbr_row[i];
        for (uint32_t i = tid; i < nr; i += PROWS_SWEEP_THREADS) {
            s_ord[i] = vg.row_ord[i];
            s_crow[i] = vg.class_rows[i];
            s_a[i] = vg.row_a[i];
        }
        v.nbr_row = s_nbr; v.row_ord = s_ord; v.class_rows = s_crow; v.row_a = s_a;
    }
    __syncthreads();
    // the thread's words of the passes over whole systems (energies, pair dots): i = tid + k * threads, row / set / word of the first
    // one computed once; when 2 W divides the thread count the set and word stay and the row advances by a constant
    const uint32_t W2 = 2u * (uint32_t)W, w_r0 = (uint32_t)tid / W2, w_p = ((uint32_t)tid / (uint32_t)W) & 1u, w_w = (uint32_t)tid % (uint32_t)W;
    const bool w_regular = PROWS_SWEEP_THREADS % W2 == 0;
    const uint32_t w_rstep = PROWS_SWEEP_THREADS / W2;
    uint32_t T[NM];
#pragma unroll
    for (int u = 0; u < NM; u++) T[u] = SHIFTED ? thr_sm[u] << 8 : thr_sm[u];
    const uint64_t key = v.keys[d];
    const PhiloxKeys ks = philox_keys((uint32_t)key, (uint32_t)(key >> 32));
    int dls[Z];
#pragma unroll
    for (int k = 0; k < Z; k++) dls[k] = v.dl[k];
    // energy.rs:99-108: every bond once through its forward direction; down spins of both sets.  Integer partial sums: warp
    // REDUX, shared-memory atomics, one barrier (cnt_sm is left zeroed for the next use).
    auto energy_phase = [&](const bool mags, const int fb, const int fs) {  // fb / fs: batch buffer and slot of a recorded sweep
        int unsat = 0, dn = 0;
        const uint32_t n_words = (uint32_t)pv.sys_words;
        uint32_t rr = w_r0;
        for (uint32_t i = tid; i < n_words; i += PROWS_SWEEP_THREADS, rr += w_rstep) {
            const uint32_t r = w_regular ? rr : i / W2;
            const int p = w_regular ? (int)w_p : (int)((i / (uint32_t)W) & 1u), w = w_regular ? (int)w_w : (int)(i % (uint32_t)W);
            const uint32_t C = prows_sm[i];
            dn += __popc(C);
#pragma unroll
            for (int k = 0; k < Z; k++) unsat += __popc(C ^ prows_nbr_word(prows_sm, W, v.nbr_row[((size_t)r * Z + k) * 2], p, w, dls[k]));
        }
        unsat = __reduce_add_sync(0xFFFFFFFFu, unsat);
        dn = __reduce_add_sync(0xFFFFFFFFu, dn);
        if ((tid & 31) == 0) {
            atomicAdd(&cnt_sm[0], (unsigned int)unsat);
            atomicAdd(&cnt_sm[1], (unsigned int)dn);
        }
        __syncthreads();
        if (tid == 0) {
            const long long tu = (long long)cnt_sm[0], td = (long long)cnt_sm[1];
            cnt_sm[0] = cnt_sm[1] = 0u;
            const float e = __fdiv_rn((float)((long long)Z * m.N - 2 * tu), (float)m.N);
            m.energies[sysg] = e;
            if (mags) m.mags[sysg] = m.N - 2 * td;
            if (cl.on) {  // rank 0 folds: hand it the scalars (visible to it after the next cluster barrier)
                cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
                *cluster.map_shared_rank(&all_e[fb][fs][rank], 0) = e;
                *cluster.map_shared_rank(&all_m[fb][fs][rank], 0) = m.N - 2 * td;
            }
        }
    };
    double *my_sums = nullptr;  // cluster mode, rank 0: the 11 running sums of (d, t)
    if (cl.on && rank == 0) {
        for (int i = tid; i < 11; i += PROWS_SWEEP_THREADS) sums_sm[i] = cl.st.sums[d * 11 * m.T + t + (int64_t)i * m.T];
        my_sums = sums_sm;
    }

#ifdef PP_PROWS_TIMING
    unsigned long long clk_last_ = clock64();
#endif
    for (int sw = 0; sw < n_sweeps; sw++) {
        PROWS_CLK(7);
        for (int colour = 0; colour < m.n_colours; colour++) {
            const int cls = colour % v.m_half;
            const uint32_t row0 = v.class_start[cls], n_items = (v.class_start[cls + 1] - row0) * (uint32_t)W;
            const uint32_t tagc = TAG_SWEEP_PACKED | (uint32_t)colour;
            for (uint32_t ih = tid; ih < 2u * n_items; ih += PROWS_SWEEP_THREADS) {
                const uint32_t it = ih >> 1;
                const int h = (int)(ih & 1u);  // sites 16 h .. 16 h + 15 of the word
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
                uint32_t M[BS ? 3 : NM];  // BS: the bit planes of L
#pragma unroll
                for (int u = 0; u < (BS ? 3 : NM); u++) M[u] = 0u;
                const uint32_t q = v.row_ord[r] * (uint32_t)W + (uint32_t)w;  // rank >> 5 of the word's sites
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
                        const uint32_t bit = 1u << (4 * g + j);  // inside the half; moved to 16 h + . below
                        const uint32_t y = SHIFTED ? ys[j] : ys[j] >> 8;
                        if (BS) {
                            const bool p1 = y < T[3];
                            const bool p2 = y < (p1 ? T[1] : T[5]);
                            const bool p3 = y < (p1 ? (p2 ? T[0] : T[2]) : (p2 ? T[4] : T[6]));
                            if (p1) M[2] |= bit;  // L = 4 p1 + 2 p2 + p3 thresholds lie above the draw
                            if (p2) M[1] |= bit;
                            if (p3) M[0] |= bit;
                        } else {
#pragma unroll
                            for (int u = 0; u < NM; u++)
                                if (y < T[u]) M[u] |= bit;
                        }
                    }
                }
                const int hs = 16 * h;
                uint32_t flip = 0u;
                if (BS) {  // unsat + L + 1 >= 8: the carry out of a three-bit add with carry in
                    const uint32_t c0 = un[0] | (M[0] << hs);
                    const uint32_t l1 = M[1] << hs, l2 = M[2] << hs;
                    const uint32_t c1 = (un[1] & l1) | (c0 & (un[1] | l1));
                    flip = (un[2] & l2) | (c1 & (un[2] | l2));
                } else {
                    // flip = OR_u ([unsat == u] & M_u); unsat >= NM always flips when NM == Z (Metropolis, energy change <= 0)
#pragma unroll
                    for (int u = 0; u <= Z2; u++) {
                        const uint32_t eq = ((u & 1) ? un[0] : ~un[0]) & ((u & 2) ? un[1] : ~un[1]) & ((u & 4) ? un[2] : ~un[2]) & ((u & 8) ? un[3] : ~un[3]);
                        flip |= u < NM ? (eq & (M[u] << hs)) : eq;
                    }
                }
                reinterpret_cast<uint16_t *>(slf)[h] = (uint16_t)((C ^ flip) >> hs);  // this thread's half of the word
            }
            __syncthreads();
        }
        const bool rec = cl.on && sw >= cl.rec_from, last = sw == n_sweeps - 1;
        PROWS_CLK(0);
        const int n_rec = rec ? sw - cl.rec_from : 0, fb = (n_rec / FOLD_K) & 1, fs = n_rec % FOLD_K;
        if (rec || (last && want_energy)) energy_phase(rec || want_mags != 0, fb, fs);
        PROWS_CLK(1);
        if (rec) {  // mod.rs:527-529, 543-578 with this launch's (fixed) system_ids
            namespace cg = cooperative_groups;
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();  // every replica of the slot has finished the sweep; its words and scalars are final
            PROWS_CLK(2);
            if (m.P > 0 && (rank & 1) == 0 && rank + 1 < m.R) {  // overlap.rs:259-281 for the pair (rank, rank + 1)
                const uint32_t n_words = (uint32_t)pv.sys_words;
                uint32_t *X = prows_sm + n_words;  // bit set where the two replicas differ
                const uint32_t *other = cluster.map_shared_rank(prows_sm, rank + 1);
                for (uint32_t i = tid; i < n_words; i += PROWS_SWEEP_THREADS) X[i] = prows_sm[i] ^ other[i];
                __syncthreads();
                int neg_q = 0, neg_l = 0;
                uint32_t rr = w_r0;
                for (uint32_t i = tid; i < n_words; i += PROWS_SWEEP_THREADS, rr += w_rstep) {
                    const uint32_t r = w_regular ? rr : i / W2;
                    const int p = w_regular ? (int)w_p : (int)((i / (uint32_t)W) & 1u), w = w_regular ? (int)w_w : (int)(i % (uint32_t)W);
                    const uint32_t x = X[i];
                    neg_q += __popc(x);
#pragma unroll
                    for (int k = 0; k < Z; k++) neg_l += __popc(x ^ prows_nbr_word(X, W, v.nbr_row[((size_t)r * Z + k) * 2], p, w, dls[k]));
                }
                neg_q = __reduce_add_sync(0xFFFFFFFFu, neg_q);
                neg_l = __reduce_add_sync(0xFFFFFFFFu, neg_l);
                if ((tid & 31) == 0) {
                    atomicAdd(&cnt_sm[2], (unsigned int)neg_q);
                    atomicAdd(&cnt_sm[3], (unsigned int)neg_l);
                }
                __syncthreads();
                if (tid == 0) {
                    const long long tq = (long long)cnt_sm[2], tl = (long long)cnt_sm[3];
                    cnt_sm[2] = cnt_sm[3] = 0u;
                    const int64_t idx = (d * m.P + rank / 2) * m.T + t;
                    const long long dq = m.N - 2 * tq, dl = (long long)Z * m.N - 2 * tl;
                    cl.dot_spin[idx] = dq;
                    cl.dot_link[idx] = dl;
                    *cluster.map_shared_rank(&all_q[fb][fs][rank / 2], 0) = dq;
                    *cluster.map_shared_rank(&all_l[fb][fs][rank / 2], 0) = dl;
                }
            }
            PROWS_CLK(3);
            cluster.sync();  // the partner's words have been read (the next sweep may change them); dots and scalars are visible
            PROWS_CLK(4);
            if (rank == 0 && tid < 32 && (fs == FOLD_K - 1 || last))  // warp 0 folds a batch (the rest of the CTA is in the next sweep)
                fold_batch(m, cl.st, d, t, tid, fs + 1, all_m[fb], all_e[fb], all_q[fb], all_l[fb], fold_tab, my_sums);
            PROWS_CLK(5);
        }
    }
    if (cl.on && n_sweeps > 0 && n_sweeps - 1 >= cl.rec_from) {
        cooperative_groups::this_cluster().sync();  // no CTA leaves while another may still address its shared memory
        if (rank == 0 && tid == 0)
            for (int i = 0; i < 11; i++) cl.st.sums[d * 11 * m.T + t + (int64_t)i * m.T] = sums_sm[i];
    }
    if (want_energy && n_sweeps == 0) energy_phase(want_mags != 0, 0, 0);
    if (n_sweeps > 0)
        for (int64_t i = tid; i < pv.sys_words / 4; i += PROWS_SWEEP_THREADS) reinterpret_cast<uint4 *>(gw)[i] = reinterpret_cast<const uint4 *>(prows_sm)[i];
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


// ------------------------------------------------------------------------------------------------
// Small ferromagnetic realizations (README quickstart sizes: 32 systems of 32 x 32 sites): the bit-packed form of
// rows_resident_kernel.  ONE CTA owns realization d for up to 256 sweeps per launch and runs the whole per-sweep sequence of
// simulation/mod.rs:405-432, 486-509, 527-529, 748-796 in shared memory; the global state stays int8 (every other kernel and the
// API keep working on it), the CTA packs it on the way in and unpacks it on the way out.
//
// Packing here is by FULL ROWS: word w of row r of a system = its sites x = 32 w .. 32 w + 31 (last extent a multiple of 32), so
// the neighbour along the row is a one-bit funnel shift with the adjacent word (a rotate when the row is one word), the neighbours
// in the other directions are the same word of row nbr_row[r][k] (shifted too when the offset's last component is +-1), and a
// colour pass updates the sites under the row's parity mask 0x5555... << p.  The 16 active sites of a word are 16 consecutive
// ranks of their colour class = half a block of the packed draw mapping: three Philox calls.  Energies and overlap dots are
// popcounts over the packed words.  Measured at the quickstart (profiles/r2_summary.md): the int8 resident kernel issues 90
// instructions per attempt (92 k warp instructions per sweep on its one SM: 19.6 us per sweep); this form runs a sweep in 8.5 us
// (pure sweeps 4.9, exchange + 2.1, measurements + 3.0).  Prefetching the exchange's log-table lines under the colour passes
// and 1024 threads per CTA were measured: no gain / 7 % slower.
constexpr int PRES_THREADS = 512;  // measured at the quickstart: 512 threads 42.3 ms per 5000 sweeps, 1024 threads 45.5 ms

template <int Z, int NM>
__global__ void __launch_bounds__(PRES_THREADS)
prows_resident_kernel(ModelView mg, RowsView vg, StatsView stg, PtView ptg, ResidentArgs a) {
    extern __shared__ __align__(16) uint32_t res_sm[];
    constexpr int Z2 = 2 * Z, NTH = 2 * Z + 1, RESIDENT_THREADS = PRES_THREADS;  // (shadows the int8 kernel's block size below)
    const int tid = threadIdx.x;
    const int64_t dg = blockIdx.x;
    const int S = mg.S, T = mg.T, P = mg.P, n_edges = T > 1 ? T - 1 : 1, L = vg.L, W = L / 32;
    const int n_rows = (int)vg.n_rows, sysw = n_rows * W;  // words per system
    uint32_t *thr_sm = res_sm;  // [T][NTH] acceptance counts (NM == Z: already shifted left by 8 for the compared ones)
    unsigned char *sc = reinterpret_cast<unsigned char *>(res_sm + ((T * NTH + 3) & ~3));
    double *sums_sm = reinterpret_cast<double *>(sc);
    long long *mag_sm = reinterpret_cast<long long *>(sums_sm + 11 * T);
    long long *dsp_sm = mag_sm + S, *dlk_sm = dsp_sm + P * T;
    unsigned long long *ea_sm = reinterpret_cast<unsigned long long *>(dlk_sm + P * T), *eacc_sm = ea_sm + n_edges, *rt_sm = eacc_sm + n_edges;
    float *en_sm = reinterpret_cast<float *>(rt_sm + S), *temps_sm = en_sm + S;
    int32_t *sid_sm = reinterpret_cast<int32_t *>(temps_sm + T);
    uint8_t *trip_sm = reinterpret_cast<uint8_t *>(sid_sm + S);
    uint32_t *pw = reinterpret_cast<uint32_t *>(sc + resident_scalar_bytes(S, T, P));  // [S][rows][W] by system
    int *cnt_sm = reinterpret_cast<int *>(pw + (size_t)S * sysw);                       // [2 * max(S, P * T)] integer partial sums
    float *dbeta_sm = reinterpret_cast<float *>(cnt_sm + 2 * (S > P * T ? S : P * T));  // [T - 1] 1 / T_e - 1 / T_e+1 (tempering.rs:85-88)
    int8_t *g_spins = mg.spins + dg * S * mg.N;
    const int64_t bins = mg.N + 1;
    for (int i = tid; i < T * NTH; i += RESIDENT_THREADS) {
        const int u = i % NTH;
        const uint32_t c = mg.lut[(i / NTH) * (4 * Z + 1) + 2 * u];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
        thr_sm[i] = (NM == Z && u < Z) ? c << 8 : c;
    }
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
    for (int i = tid; i < T - 1; i += RESIDENT_THREADS) dbeta_sm[i] = __fsub_rn(__fdiv_rn(1.0f, mg.temps[i]), __fdiv_rn(1.0f, mg.temps[i + 1]));
    for (int i = tid; i < S * sysw; i += RESIDENT_THREADS) {  // pack: 32 consecutive sites -> one word (bit = 1: spin -1)
        const uint4 *src = reinterpret_cast<const uint4 *>(g_spins + (int64_t)i * 32);
        uint32_t word = 0u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 q4 = src[h];
            const uint32_t part[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t sg = part[j] & 0x80808080u;  // the sign bits of four int8 spins
                word |= (((sg >> 7) & 1u) | ((sg >> 14) & 2u) | ((sg >> 21) & 4u) | ((sg >> 28) & 8u)) << (16 * h + 4 * j);
            }
        }
        pw[i] = word;
    }
    ModelView m = mg;
    m.D = 1;
    m.sample_offset = mg.sample_offset + dg;  // seeds stay those of the global realization index
    m.system_ids = sid_sm;
    m.energies = en_sm;
    m.mags = mag_sm;
    m.temps = temps_sm;
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
    const uint64_t key = vg.keys[dg];
    const PhiloxKeys ks = philox_keys((uint32_t)key, (uint32_t)(key >> 32));
    int dls[Z];
#pragma unroll
    for (int k = 0; k < Z; k++) dls[k] = vg.dl[k];
    // word w of row nr of the system at `sys`, as seen by the sites of word w after moving `sh` sites along the row
    auto nbr_word = [&](const uint32_t *sys, const uint32_t nr, const int w, const int sh) {
        const uint32_t *row = sys + (size_t)nr * W;
        const uint32_t c = row[w];
        if (sh == 0) return c;
        if (sh > 0) return (c >> 1) | (row[w + 1 == W ? 0 : w + 1] << 31);
        return (c << 1) | (row[w ? w - 1 : W - 1] >> 31);
    };
    __syncthreads();
    uint32_t pt_event = a.pt_event0;
    int parity = a.parity0;
    const int half_l = L / 2;
    // one word of one system: thresholds against the packed draws of its 16 active sites, flip, store
    auto update_word = [&](const int slot, const int sysl, uint32_t *own, const uint32_t C, const uint32_t (&b)[Z2], const int p,
                           const uint32_t q, const uint32_t h, const uint32_t sweep_index, const uint32_t tagc) {
        const uint32_t *thr = thr_sm + (slot % T) * NTH;      // realization.rs:166
        uint32_t un[4];
        prows_count<Z2>(b, un);
        uint32_t Tm[NM], M[NM];
#pragma unroll
        for (int u = 0; u < NM; u++) { Tm[u] = thr[u]; M[u] = 0u; }
        uint32_t wd[12];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const u32x4 o = philox4x32_k(q, sweep_index, (uint32_t)sysl, tagc | ((3u * h + (uint32_t)c) << 8), ks);
            wd[4 * c] = o.x; wd[4 * c + 1] = o.y; wd[4 * c + 2] = o.z; wd[4 * c + 3] = o.w;
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const uint32_t A = wd[3 * g], B = wd[3 * g + 1], Cw = wd[3 * g + 2];
            const uint32_t y3 = __byte_perm(__byte_perm(Cw, B, 0x0400), A, 0x4210);
            const uint32_t ys[4] = {A, B, Cw, y3};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t bit = 1u << (2 * (4 * g + j));  // site x = 2 (4 g + j) + p: shifted by p below
                const uint32_t y = NM == Z ? ys[j] : ys[j] >> 8;
#pragma unroll
                for (int u = 0; u < NM; u++)
                    if (y < Tm[u]) M[u] |= bit;
            }
        }
        uint32_t flip = 0u;
#pragma unroll
        for (int u = 0; u <= Z2; u++) {
            const uint32_t eq = ((u & 1) ? un[0] : ~un[0]) & ((u & 2) ? un[1] : ~un[1]) & ((u & 4) ? un[2] : ~un[2]) & ((u & 8) ? un[3] : ~un[3]);
            flip |= u < NM ? (eq & (M[u] << p)) : eq;
        }
        *own = C ^ (flip & (0x55555555u << p));
    };
    // Two-colour lattices whose work items fit two per thread (the quickstart: 32 systems x 32 row words on 512 threads): a thread's
    // items are the same words in every colour pass of every sweep (only the parity of the active sites alternates), so their slot,
    // word offsets, neighbour-word offsets and rank block are worked out ONCE per launch instead of once per pass (two integer
    // divisions and nine dependent table reads per item: the passes of one small realization on one SM are latency-bound).
    constexpr int HOIST = 2;
    const uint32_t per_sys0 = (vg.class_start[1] - vg.class_start[0]) * (uint32_t)W;
    const bool hoist = m.n_colours == 2 && vg.m_half == 1 && (uint32_t)S * per_sys0 <= (uint32_t)(HOIST * RESIDENT_THREADS) && sysw <= 65535;
    int h_slot[HOIST], h_a[HOIST];
    uint32_t h_own[HOIST], h_q[HOIST], h_h[HOIST];
    uint16_t h_n0[HOIST][Z2], h_n1[HOIST][Z2];  // word offsets inside a system: the neighbour row's word, and the adjacent one for a shift
#pragma unroll
    for (int j = 0; j < HOIST; j++) {
        h_slot[j] = -1;
        const uint32_t it = (uint32_t)tid + (uint32_t)j * RESIDENT_THREADS;
        if (hoist && it < (uint32_t)S * per_sys0) {
            const int slot = (int)(it / per_sys0);
            const uint32_t rem = it - (uint32_t)slot * per_sys0, ri = rem / (uint32_t)W;
            const int w = (int)(rem - ri * (uint32_t)W);
            const uint32_t r = vg.class_rows[ri];
            h_slot[j] = slot;
            h_a[j] = (int)vg.row_a[r];
            h_own[j] = r * (uint32_t)W + (uint32_t)w;
            const uint32_t base = vg.row_ord[r] * (uint32_t)half_l + 16u * (uint32_t)w;
            h_q[j] = base >> 5;
            h_h[j] = (base >> 4) & 1u;
#pragma unroll
            for (int k = 0; k < Z; k++)
#pragma unroll
                for (int sgn = 0; sgn < 2; sgn++) {
                    const uint32_t nr = vg.nbr_row[((size_t)r * Z + k) * 2 + sgn];
                    const int sh = sgn ? -dls[k] : dls[k];
                    const int wa = sh > 0 ? (w + 1 == W ? 0 : w + 1) : sh < 0 ? (w ? w - 1 : W - 1) : w;
                    h_n0[j][2 * k + sgn] = (uint16_t)(nr * (uint32_t)W + (uint32_t)w);
                    h_n1[j][2 * k + sgn] = (uint16_t)(nr * (uint32_t)W + (uint32_t)wa);
                }
        }
    }
    for (int sw = 0; sw < a.n_sweeps; sw++) {
        const int64_t sid = a.sweep_id0 + sw;
        const uint32_t sweep_index = a.sweep_counter0 + (uint32_t)sw;
        // single-edge exchange after this sweep: its edge and draw depend on (event, replica) only — generate them now and start the
        // read of ln u (a 64 MiB table: a DRAM round trip) so that the exchange itself is a compare on values already here
        int pt_edge = 0;
        float pt_logu = 0.0f;
        if (a.pt_schedule == 0 && T >= 2 && tid < m.R && a.pt_interval > 0 && sid % a.pt_interval == 0) {
            const u32x4 o = philox4x32(0xFFFFFFFFu, pt_event, (uint32_t)tid, TAG_PT, (uint32_t)key, (uint32_t)(key >> 32));
            pt_edge = (int)(((uint64_t)o.y * (uint64_t)(T - 1)) >> 32);
            pt_logu = m.logtab[o.x >> 8];
        }
        for (int colour = 0; colour < m.n_colours; colour++) {
            const uint32_t tagc = TAG_SWEEP_PACKED | (uint32_t)colour;
            if (hoist) {
#pragma unroll
                for (int j = 0; j < HOIST; j++) {
                    if (h_slot[j] < 0) continue;
                    const int sysl = sid_sm[h_slot[j]];  // parallel.rs:27-33
                    uint32_t *sys = pw + (size_t)sysl * sysw;
                    const uint32_t C = sys[h_own[j]];
                    uint32_t b[Z2];
#pragma unroll
                    for (int k = 0; k < Z; k++)
#pragma unroll
                        for (int sgn = 0; sgn < 2; sgn++) {
                            const int sh = sgn ? -dls[k] : dls[k];
                            const uint32_t c0 = sys[h_n0[j][2 * k + sgn]];
                            uint32_t nw = c0;
                            if (sh > 0) nw = (c0 >> 1) | (sys[h_n1[j][2 * k + sgn]] << 31);
                            else if (sh < 0) nw = (c0 << 1) | (sys[h_n1[j][2 * k + sgn]] >> 31);
                            b[2 * k + sgn] = C ^ nw;
                        }
                    update_word(h_slot[j], sysl, sys + h_own[j], C, b, h_a[j] == colour ? 0 : 1, h_q[j], h_h[j], sweep_index, tagc);
                }
                __syncthreads();
                continue;
            }
            const int cls = colour % vg.m_half;
            const uint32_t row0 = vg.class_start[cls], per_sys = (vg.class_start[cls + 1] - row0) * (uint32_t)W;
            for (uint32_t it = tid; it < (uint32_t)S * per_sys; it += RESIDENT_THREADS) {
                const int slot = (int)(it / per_sys);
                const uint32_t rem = it - (uint32_t)slot * per_sys, ri = rem / (uint32_t)W;
                const int w = (int)(rem - ri * (uint32_t)W);
                const uint32_t r = vg.class_rows[row0 + ri];
                const int p = (int)vg.row_a[r] == colour ? 0 : 1;  // parity of the row's sites of this colour
                const int sysl = sid_sm[slot];                        // parallel.rs:27-33
                uint32_t *sys = pw + (size_t)sysl * sysw;
                const uint32_t C = sys[(size_t)r * W + w];
                uint32_t b[Z2];
#pragma unroll
                for (int k = 0; k < Z; k++) {
                    b[2 * k] = C ^ nbr_word(sys, vg.nbr_row[((size_t)r * Z + k) * 2], w, dls[k]);
                    b[2 * k + 1] = C ^ nbr_word(sys, vg.nbr_row[((size_t)r * Z + k) * 2 + 1], w, -dls[k]);
                }
                // the word's 16 sites of this colour are the ranks base .. base + 15: half h of block q of the packed mapping
                const uint32_t base = vg.row_ord[r] * (uint32_t)half_l + 16u * (uint32_t)w;
                update_word(slot, sysl, sys + (size_t)r * W + w, C, b, p, base >> 5, (base >> 4) & 1u, sweep_index, tagc);
            }
            __syncthreads();
        }
        const bool record = sid >= a.warmup_sweeps;
        const bool pt_this = a.pt_interval > 0 && sid % a.pt_interval == 0;
        if (record || pt_this) {  // mod.rs:486-509; energy.rs:99-108: every bond once through its forward direction
            for (int i = tid; i < 2 * S; i += RESIDENT_THREADS) cnt_sm[i] = 0;
            __syncthreads();
            for (int i0 = (tid & ~31); i0 < S * sysw; i0 += RESIDENT_THREADS) {  // whole warps: the sums of a warp that sits
                const int i = i0 + (tid & 31);                                   // inside one system meet in one reduction
                const bool live = i < S * sysw;
                const int sysl = live ? i / sysw : -1;
                int unsat = 0, dn = 0;
                if (live) {
                    const int rem = i - sysl * sysw, r = rem / W, w = rem - r * W;
                    const uint32_t *sys = pw + (size_t)sysl * sysw;
                    const uint32_t C = sys[rem];
#pragma unroll
                    for (int k = 0; k < Z; k++) unsat += __popc(C ^ nbr_word(sys, vg.nbr_row[((size_t)r * Z + k) * 2], w, dls[k]));
                    dn = __popc(C);
                }
                if (__all_sync(0xFFFFFFFFu, sysl == __shfl_sync(0xFFFFFFFFu, sysl, 0))) {
                    unsat = __reduce_add_sync(0xFFFFFFFFu, unsat);
                    dn = __reduce_add_sync(0xFFFFFFFFu, dn);
                    if ((tid & 31) == 0 && live) { atomicAdd(&cnt_sm[2 * sysl], unsat); atomicAdd(&cnt_sm[2 * sysl + 1], dn); }
                } else if (live) {
                    atomicAdd(&cnt_sm[2 * sysl], unsat);
                    atomicAdd(&cnt_sm[2 * sysl + 1], dn);
                }
            }
            __syncthreads();
            for (int sysl = tid; sysl < S; sysl += RESIDENT_THREADS) {
                en_sm[sysl] = __fdiv_rn((float)((long long)Z * m.N - 2ll * cnt_sm[2 * sysl]), (float)m.N);
                if (record) mag_sm[sysl] = m.N - 2ll * cnt_sm[2 * sysl + 1];
            }
            __syncthreads();
        }
        if (record) {
            if (P > 0) {  // overlap.rs:259-281 with this sweep's pre-exchange system_ids
                for (int i = tid; i < 2 * P * T; i += RESIDENT_THREADS) cnt_sm[i] = 0;
                __syncthreads();
                for (int i0 = (tid & ~31); i0 < P * T * sysw; i0 += RESIDENT_THREADS) {
                    const int i = i0 + (tid & 31);
                    const bool live = i < P * T * sysw;
                    const int idx = live ? i / sysw : -1;
                    int neg_q = 0, neg_l = 0;
                    if (live) {
                        const int rem = i - idx * sysw, r = rem / W, w = rem - r * W;
                        const int t = idx % T, pr = idx / T;
                        const uint32_t *sa = pw + (size_t)sid_sm[(2 * pr) * T + t] * sysw, *sb = pw + (size_t)sid_sm[(2 * pr + 1) * T + t] * sysw;
                        const uint32_t x = sa[rem] ^ sb[rem];
#pragma unroll
                        for (int k = 0; k < Z; k++) {
                            const uint32_t nr = vg.nbr_row[((size_t)r * Z + k) * 2];
                            neg_l += __popc(x ^ nbr_word(sa, nr, w, dls[k]) ^ nbr_word(sb, nr, w, dls[k]));
                        }
                        neg_q = __popc(x);
                    }
                    if (__all_sync(0xFFFFFFFFu, idx == __shfl_sync(0xFFFFFFFFu, idx, 0))) {
                        neg_q = __reduce_add_sync(0xFFFFFFFFu, neg_q);
                        neg_l = __reduce_add_sync(0xFFFFFFFFu, neg_l);
                        if ((tid & 31) == 0 && live) { atomicAdd(&cnt_sm[2 * idx], neg_q); atomicAdd(&cnt_sm[2 * idx + 1], neg_l); }
                    } else if (live) {
                        atomicAdd(&cnt_sm[2 * idx], neg_q);
                        atomicAdd(&cnt_sm[2 * idx + 1], neg_l);
                    }
                }
                __syncthreads();
                for (int idx = tid; idx < P * T; idx += RESIDENT_THREADS) {
                    dsp_sm[idx] = m.N - 2ll * cnt_sm[2 * idx];
                    dlk_sm[idx] = (long long)Z * m.N - 2ll * cnt_sm[2 * idx + 1];
                }
                __syncthreads();
            }
            for (int idx = tid; idx < 12 * T; idx += RESIDENT_THREADS) {  // mod.rs:543-578: one thread per (temperature, running sum)
                const int t = idx / 12;
                fold_spread<true>(
                    m, st, 0, t, idx - 12 * t,
                    [&](int r) { return mag_sm[sid_sm[r * T + t]]; },
                    [&](int r) { return en_sm[sid_sm[r * T + t]]; },
                    [&](int p) { return dsp_sm[p * T + t]; },
                    [&](int p) { return dlk_sm[p * T + t]; });
            }
            __syncthreads();
        }
        if (pt_this) {  // mod.rs:748-796
            if (T >= 2) {
                if (a.pt_schedule == 0) {  // tempering.rs:20-42: edge, draw and ln u were prepared while the sweep ran
                    if (tid < m.R) pt_attempt_edge_pre(m, pt, 0, tid, pt_edge, pt_logu, dbeta_sm[pt_edge]);
                } else {
                    if (tid < m.R) pt_exchange_body(m, pt, 0, tid, a.pt_schedule, parity, pt_event);
                    parity = 1 - parity;
                }
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
        const_cast<long long *>(stg.dot_spin)[dg * P * T + i] = dsp_sm[i];
        const_cast<long long *>(stg.dot_link)[dg * P * T + i] = dlk_sm[i];
    }
    for (int i = tid; i < T - 1; i += RESIDENT_THREADS) {
        ptg.edge_attempts[dg * (T - 1) + i] = ea_sm[i];
        ptg.edge_acceptances[dg * (T - 1) + i] = eacc_sm[i];
    }
    for (int i = tid; i < S * sysw; i += RESIDENT_THREADS) {  // unpack: +1 = 0x01, -1 = 0xFF
        const uint32_t word = pw[i];
        uint4 *dst = reinterpret_cast<uint4 *>(g_spins + (int64_t)i * 32);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t part[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t nib = (word >> (16 * h + 4 * j)) & 15u;
                const uint32_t spread = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);  // one bit per byte
                part[j] = 0x01010101u ^ (spread * 0xFEu);  // 0x01 ^ 0xFE = 0xFF where the bit is set
            }
            dst[h] = make_uint4(part[0], part[1], part[2], part[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same small realizations spread over a thread-block CLUSTER (round 2; the one-CTA form above is bound by the instruction
// throughput of its one SM: 13 instructions per attempt x 32 Ki attempts per sweep).  NC CTAs per realization, CTA c keeps the words
// of the systems c S/NC .. (c + 1) S/NC - 1 in its shared memory for the whole launch; what crosses CTAs goes through distributed
// shared memory:
//   * after the colour passes a CTA counts bonds / down spins of its systems and stores energy (+ magnetisation sum) of each into the
//     `en_all` / `mag_all` arrays of EVERY CTA of the cluster; one cluster barrier later every CTA holds all S values;
//   * every CTA then replays the exchange on its own copy of system_ids / counters (same draws, same energies: identical decisions,
//     tempering.rs:20-102) — no broadcast of the result is needed; rank 0's copy is written back at the end;
//   * recorded sweeps: the pair (replica 2p, replica 2p + 1) at slot t is handled by a warp of the CTA that owns the first system,
//     reading the partner's words through DSMEM (overlap.rs:259-281); the dots go to the CTA that folds temperature t (t mod NC), which
//     keeps the 11 running sums of its temperatures in shared memory for the whole launch (mod.rs:543-578 as fold_spread);
//   * one cluster barrier per sweep with an exchange, two per recorded sweep; everything that is stored remotely is double-buffered
//     by sweep parity (a CTA may be most of a sweep ahead of another).
// Requirements (host-checked): two colours, one row class (m_half = 1), S a multiple of NC, work items <= 2 per thread.
struct ClusterResidentLayout {  // word offsets into the dynamic shared memory of one CTA (all 8-byte aligned blocks first)
    size_t sums, mag_all, dsp_all, dlk_all, ea, eacc, rt, en_all, temps, dbeta, sid, cnt, thr, pw, snap, xs, trip, total_bytes;
};
__host__ __device__ inline ClusterResidentLayout cluster_resident_layout(int S, int T, int P, int z, int nc, int sysw) {
    ClusterResidentLayout l;
    const size_t n_edges = T > 1 ? (size_t)(T - 1) : 1;
    size_t o = 0;  // in bytes
    l.sums = o; o += 8 * (size_t)11 * T;
    l.mag_all = o; o += 8 * 2 * (size_t)S;
    l.dsp_all = o; o += 8 * 2 * (size_t)(P * T > 0 ? P * T : 1);
    l.dlk_all = o; o += 8 * 2 * (size_t)(P * T > 0 ? P * T : 1);
    l.ea = o; o += 8 * n_edges;
    l.eacc = o; o += 8 * n_edges;
    l.rt = o; o += 8 * (size_t)S;
    l.en_all = o; o += 4 * 2 * (size_t)S;
    l.temps = o; o += 4 * (size_t)T;
    l.dbeta = o; o += 4 * n_edges;
    l.sid = o; o += 4 * (size_t)S;
    l.cnt = o; o += 4 * 2 * (size_t)(S / nc);
    l.thr = o; o += 4 * (size_t)T * (2 * z + 1);
    o = (o + 15) & ~size_t(15);
    l.pw = o; o += 4 * (size_t)(S / nc) * sysw;
    l.snap = o; o += 4 * 2 * (size_t)(S / nc) * sysw;  // the words as the last two recorded sweeps left them (read by the pair dots)
    l.xs = o; o += 4 * (size_t)16 * sysw;  // one XOR-word scratch per warp (<= 16 warps)
    l.trip = o; o += (size_t)S;
    l.total_bytes = (o + 15) & ~size_t(15);
    return l;
}

template <int Z, int NM>
__global__ void __launch_bounds__(512)
prows_cluster_resident_kernel(ModelView mg, RowsView vg, StatsView stg, PtView ptg, ResidentArgs a, int NC) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char cres_sm[];
    constexpr int Z2 = 2 * Z, NTH = 2 * Z + 1;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, n_warps = NT >> 5;
    const int rank = (int)(blockIdx.x % NC);
    const int64_t dg = blockIdx.x / NC;
    const int S = mg.S, T = mg.T, P = mg.P, L = vg.L, W = L / 32;
    const int n_rows = (int)vg.n_rows, sysw = n_rows * W, SPC = S / NC, sys0 = rank * SPC;  // this CTA owns systems sys0 .. sys0 + SPC - 1
    const ClusterResidentLayout lay = cluster_resident_layout(S, T, P, Z, NC, sysw);
    double *sums_sm = reinterpret_cast<double *>(cres_sm + lay.sums);
    long long *mag_all = reinterpret_cast<long long *>(cres_sm + lay.mag_all);  // [2][S] by sweep parity
    long long *dsp_all = reinterpret_cast<long long *>(cres_sm + lay.dsp_all);  // [2][P * T]
    long long *dlk_all = reinterpret_cast<long long *>(cres_sm + lay.dlk_all);
    unsigned long long *ea_sm = reinterpret_cast<unsigned long long *>(cres_sm + lay.ea);
    unsigned long long *eacc_sm = reinterpret_cast<unsigned long long *>(cres_sm + lay.eacc);
    unsigned long long *rt_sm = reinterpret_cast<unsigned long long *>(cres_sm + lay.rt);
    float *en_all = reinterpret_cast<float *>(cres_sm + lay.en_all);            // [2][S]
    float *temps_sm = reinterpret_cast<float *>(cres_sm + lay.temps), *dbeta_sm = reinterpret_cast<float *>(cres_sm + lay.dbeta);
    int32_t *sid_sm = reinterpret_cast<int32_t *>(cres_sm + lay.sid);
    int *cnt_sm = reinterpret_cast<int *>(cres_sm + lay.cnt);                   // [2 * SPC]
    uint32_t *thr_sm = reinterpret_cast<uint32_t *>(cres_sm + lay.thr);
    uint32_t *pw = reinterpret_cast<uint32_t *>(cres_sm + lay.pw);              // [SPC][rows][W]
    uint32_t *snap_sm = reinterpret_cast<uint32_t *>(cres_sm + lay.snap);       // [2][SPC][rows][W]
    uint32_t *xs_sm = reinterpret_cast<uint32_t *>(cres_sm + lay.xs);           // [warps][rows][W]
    uint8_t *trip_sm = cres_sm + lay.trip;
    int8_t *g_spins = mg.spins + dg * S * mg.N;
    const int64_t bins = mg.N + 1;
    const int PT_ = P * T;

    for (int i = tid; i < T * NTH; i += NT) {
        const int u = i % NTH;
        const uint32_t c = mg.lut[(i / NTH) * (4 * Z + 1) + 2 * u];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
        thr_sm[i] = (NM == Z && u < Z) ? c << 8 : c;
    }
    for (int i = tid; i < 11 * T; i += NT) sums_sm[i] = stg.sums[dg * 11 * T + i];
    for (int i = tid; i < S; i += NT) {
        sid_sm[i] = mg.system_ids[dg * S + i];
        rt_sm[i] = ptg.round_trips[dg * S + i];
        trip_sm[i] = ptg.trip_state[dg * S + i];
    }
    for (int i = tid; i < T - 1; i += NT) {
        ea_sm[i] = ptg.edge_attempts[dg * (T - 1) + i];
        eacc_sm[i] = ptg.edge_acceptances[dg * (T - 1) + i];
        dbeta_sm[i] = __fsub_rn(__fdiv_rn(1.0f, mg.temps[i]), __fdiv_rn(1.0f, mg.temps[i + 1]));
    }
    for (int i = tid; i < T; i += NT) temps_sm[i] = mg.temps[i];
    for (int i = tid; i < SPC * sysw; i += NT) {  // pack the CTA's systems: 32 consecutive sites -> one word (bit = 1: spin -1)
        const uint4 *src = reinterpret_cast<const uint4 *>(g_spins + ((int64_t)sys0 * sysw + i) * 32);
        uint32_t word = 0u;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint4 q4 = src[h];
            const uint32_t part[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t sg = part[j] & 0x80808080u;
                word |= (((sg >> 7) & 1u) | ((sg >> 14) & 2u) | ((sg >> 21) & 4u) | ((sg >> 28) & 8u)) << (16 * h + 4 * j);
            }
        }
        pw[i] = word;
    }
    ModelView m = mg;
    m.D = 1;
    m.sample_offset = mg.sample_offset + dg;
    m.system_ids = sid_sm;
    m.temps = temps_sm;
    StatsView st = stg;
    st.sums = sums_sm;
    if (stg.hist) {
        st.hist = stg.hist + dg * T * bins;
        st.ql_at_q = stg.ql_at_q + dg * T * bins;
        st.ql2_at_q = stg.ql2_at_q + dg * T * bins;
    }
    PtView pt = ptg;
    pt.edge_attempts = ea_sm;
    pt.edge_acceptances = eacc_sm;
    pt.round_trips = rt_sm;
    pt.trip_state = trip_sm;
    const uint64_t key = vg.keys[dg];
    const PhiloxKeys ks = philox_keys((uint32_t)key, (uint32_t)(key >> 32));
    int dls[Z];
#pragma unroll
    for (int k = 0; k < Z; k++) dls[k] = vg.dl[k];
    auto nbr_word = [&](const uint32_t *sys, const uint32_t nr, const int w, const int sh) {
        const uint32_t *row = sys + (size_t)nr * W;
        const uint32_t c = row[w];
        if (sh == 0) return c;
        if (sh > 0) return (c >> 1) | (row[w + 1 == W ? 0 : w + 1] << 31);
        return (c << 1) | (row[w ? w - 1 : W - 1] >> 31);
    };
    const int half_l = L / 2;
    // the thread's work items (words of the CTA's systems): resolved once, the same in every colour pass (see prows_resident_kernel)
    constexpr int HOIST = 2;
    const uint32_t per_sys0 = (uint32_t)n_rows * (uint32_t)W;
    int h_ls[HOIST], h_a[HOIST];  // local system, row phase
    uint32_t h_own[HOIST], h_q[HOIST], h_h[HOIST];
    uint16_t h_n0[HOIST][Z2], h_n1[HOIST][Z2];
#pragma unroll
    for (int j = 0; j < HOIST; j++) {
        h_ls[j] = -1;
        const uint32_t it = (uint32_t)tid + (uint32_t)j * (uint32_t)NT;
        if (it < (uint32_t)SPC * per_sys0) {
            const int ls = (int)(it / per_sys0);
            const uint32_t rem = it - (uint32_t)ls * per_sys0, r = rem / (uint32_t)W;
            const int w = (int)(rem - r * (uint32_t)W);
            h_ls[j] = ls;
            h_a[j] = (int)vg.row_a[r];
            h_own[j] = rem;
            const uint32_t base = vg.row_ord[r] * (uint32_t)half_l + 16u * (uint32_t)w;
            h_q[j] = base >> 5;
            h_h[j] = (base >> 4) & 1u;
#pragma unroll
            for (int k = 0; k < Z; k++)
#pragma unroll
                for (int sgn = 0; sgn < 2; sgn++) {
                    const uint32_t nr = vg.nbr_row[((size_t)r * Z + k) * 2 + sgn];
                    const int sh = sgn ? -dls[k] : dls[k];
                    const int wa = sh > 0 ? (w + 1 == W ? 0 : w + 1) : sh < 0 ? (w ? w - 1 : W - 1) : w;
                    h_n0[j][2 * k + sgn] = (uint16_t)(nr * (uint32_t)W + (uint32_t)w);
                    h_n1[j][2 * k + sgn] = (uint16_t)(nr * (uint32_t)W + (uint32_t)wa);
                }
        }
    }
    // one work item per thread and a system per warp (the quickstart: 32 row words per system): the counts need no shared memory
    const bool warp_is_system = sysw == 32 && NT == SPC * 32 && NC <= 32;
    __syncthreads();
    // slot of each of the CTA's systems (the inverse of system_ids restricted to this CTA), rebuilt after every exchange
    __shared__ int slot_of[64];
    auto rebuild_slots = [&]() {
        for (int slot = tid; slot < S; slot += NT) {
            const int sy = sid_sm[slot] - sys0;
            if (sy >= 0 && sy < SPC) slot_of[sy] = slot;
        }
    };
    rebuild_slots();
    cluster.sync();  // every CTA's shared memory is initialised before anyone stores into it
    uint32_t pt_event = a.pt_event0;
    int parity = a.parity0;
    // single-edge exchange: edge, draw and ln u (a read of the 64 MiB table: a DRAM round trip) depend on (event, replica) only; they
    // are prepared for the NEXT event between the arrive and the wait of a cluster barrier, where the threads would only wait
    int pt_edge = 0;
    float pt_logu = 0.0f;
    const bool pt_prep = a.pt_schedule == 0 && T >= 2 && tid < m.R && a.pt_interval > 0;
    auto prepare_exchange = [&](const uint32_t event) {
        const u32x4 o = philox4x32(0xFFFFFFFFu, event, (uint32_t)tid, TAG_PT, (uint32_t)key, (uint32_t)(key >> 32));
        pt_edge = (int)(((uint64_t)o.y * (uint64_t)(T - 1)) >> 32);
        pt_logu = m.logtab[o.x >> 8];
    };
    if (pt_prep) prepare_exchange(pt_event);
    int last_e = -1, last_m = -1, last_d = -1;  // buffer of the last energies / magnetisations / pair dots of this launch
#ifdef PP_PROWS_TIMING
    unsigned long long clk_last_ = clock64();
#endif
    int sync_n = 0;  // sweeps with a cluster barrier so far: the remotely written buffers alternate with THEM (between two such
                     // sweeps no barrier keeps a CTA from running ahead)
    for (int sw = 0; sw < a.n_sweeps; sw++) {
        const int64_t sid = a.sweep_id0 + sw;
        const uint32_t sweep_index = a.sweep_counter0 + (uint32_t)sw;
        const bool record = sid >= a.warmup_sweeps;
        const bool pt_this = a.pt_interval > 0 && sid % a.pt_interval == 0;
        const int par = sync_n & 1;
        if (record || pt_this) sync_n++;
        int cur_edge = pt_edge;      // the exchange after THIS sweep (prepared before the loop or under an earlier barrier)
        float cur_logu = pt_logu;
        bool prepared_next = false;  // pt_edge / pt_logu already hold the event after this sweep's
        PROWS_CLK(7);
        // a warp that holds exactly one system counts its bonds and down spins inside the colour-1 pass (no pass of its own)
        const bool cnt_inpass = warp_is_system && (record || pt_this);
        int ip_unsat = 0, ip_dn = 0;
        for (int colour = 0; colour < 2; colour++) {
            const uint32_t tagc = TAG_SWEEP_PACKED | (uint32_t)colour;
#pragma unroll
            for (int j = 0; j < HOIST; j++) {
                if (h_ls[j] < 0) continue;
                const int slot = slot_of[h_ls[j]], sysl = sys0 + h_ls[j];
                uint32_t *sys = pw + (size_t)h_ls[j] * sysw;
                const uint32_t C = sys[h_own[j]];
                uint32_t b[Z2];
#pragma unroll
                for (int k = 0; k < Z; k++)
#pragma unroll
                    for (int sgn = 0; sgn < 2; sgn++) {
                        const int sh = sgn ? -dls[k] : dls[k];
                        const uint32_t c0 = sys[h_n0[j][2 * k + sgn]];
                        uint32_t nw = c0;
                        if (sh > 0) nw = (c0 >> 1) | (sys[h_n1[j][2 * k + sgn]] << 31);
                        else if (sh < 0) nw = (c0 << 1) | (sys[h_n1[j][2 * k + sgn]] >> 31);
                        b[2 * k + sgn] = C ^ nw;
                    }
                const int p = h_a[j] == colour ? 0 : 1;
                const uint32_t *thr = thr_sm + (slot % T) * NTH;  // realization.rs:166
                uint32_t un[4];
                prows_count<Z2>(b, un);
                uint32_t Tm[NM], M[NM];
#pragma unroll
                for (int u = 0; u < NM; u++) { Tm[u] = thr[u]; M[u] = 0u; }
                uint32_t wd[12];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const u32x4 o = philox4x32_k(h_q[j], sweep_index, (uint32_t)sysl, tagc | ((3u * h_h[j] + (uint32_t)c) << 8), ks);
                    wd[4 * c] = o.x; wd[4 * c + 1] = o.y; wd[4 * c + 2] = o.z; wd[4 * c + 3] = o.w;
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint32_t A = wd[3 * g], B = wd[3 * g + 1], Cw = wd[3 * g + 2];
                    const uint32_t y3 = __byte_perm(__byte_perm(Cw, B, 0x0400), A, 0x4210);
                    const uint32_t ys[4] = {A, B, Cw, y3};
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        const uint32_t bit = 1u << (2 * (4 * g + jj));
                        const uint32_t y = NM == Z ? ys[jj] : ys[jj] >> 8;
#pragma unroll
                        for (int u = 0; u < NM; u++)
                            if (y < Tm[u]) M[u] |= bit;
                    }
                }
                uint32_t flip = 0u;
#pragma unroll
                for (int u = 0; u <= Z2; u++) {
                    const uint32_t eq = ((u & 1) ? un[0] : ~un[0]) & ((u & 2) ? un[1] : ~un[1]) & ((u & 4) ? un[2] : ~un[2]) & ((u & 8) ? un[3] : ~un[3]);
                    flip |= u < NM ? (eq & (M[u] << p)) : eq;
                }
                const uint32_t act = 0x55555555u << p, Cn = C ^ (flip & act);
                sys[h_own[j]] = Cn;
                if (cnt_inpass && colour == 1) {  // energy.rs:99-108 from the last pass: its sites' 2z' bonds are all bonds, each once; a
                    const uint32_t fm = flip & act;  // flip inverts the bond words of its site
#pragma unroll
                    for (int k = 0; k < Z2; k++) ip_unsat += __popc((b[k] ^ fm) & act);
                    ip_dn += __popc(Cn);
                }
            }
            __syncthreads();
        }
        PROWS_CLK(0);
        if (record && P > 0)  // the pair dots read these copies: the words themselves may move on once barrier A is passed
            for (int i = tid; i < SPC * sysw; i += NT) snap_sm[(size_t)par * SPC * sysw + i] = pw[i];
        if (record || pt_this) {  // mod.rs:486-509; energy.rs:99-108 for the CTA's systems, results to every CTA of the cluster
            if (warp_is_system) {  // a warp's work items are exactly one system: totals by REDUX, lanes 0 .. NC - 1 deliver them
                const int ls = h_ls[0];
                const int unsat = __reduce_add_sync(0xFFFFFFFFu, ip_unsat);
                const int dn = __reduce_add_sync(0xFFFFFFFFu, ip_dn);
                if (lane < NC) {
                    const float e = __fdiv_rn((float)((long long)Z * m.N - 2ll * unsat), (float)m.N);
                    *cluster.map_shared_rank(&en_all[par * S + sys0 + ls], lane) = e;
                    if (record) *cluster.map_shared_rank(&mag_all[par * S + sys0 + ls], lane) = m.N - 2ll * dn;
                }
            } else {
            for (int i = tid; i < 2 * SPC; i += NT) cnt_sm[i] = 0;
            __syncthreads();
            for (int i0 = (tid & ~31); i0 < SPC * sysw; i0 += NT) {
                const int i = i0 + lane;
                const bool live = i < SPC * sysw;
                const int ls = live ? i / sysw : -1;
                int unsat = 0, dn = 0;
                if (live) {
                    const int rem = i - ls * sysw, r = rem / W, w = rem - r * W;
                    const uint32_t *sys = pw + (size_t)ls * sysw;
                    const uint32_t C = sys[rem];
#pragma unroll
                    for (int k = 0; k < Z; k++) unsat += __popc(C ^ nbr_word(sys, vg.nbr_row[((size_t)r * Z + k) * 2], w, dls[k]));
                    dn = __popc(C);
                }
                if (__all_sync(0xFFFFFFFFu, ls == __shfl_sync(0xFFFFFFFFu, ls, 0))) {
                    unsat = __reduce_add_sync(0xFFFFFFFFu, unsat);
                    dn = __reduce_add_sync(0xFFFFFFFFu, dn);
                    if (lane == 0 && live) { atomicAdd(&cnt_sm[2 * ls], unsat); atomicAdd(&cnt_sm[2 * ls + 1], dn); }
                } else if (live) {
                    atomicAdd(&cnt_sm[2 * ls], unsat);
                    atomicAdd(&cnt_sm[2 * ls + 1], dn);
                }
            }
            __syncthreads();
            for (int i = tid; i < SPC * NC; i += NT) {  // (system, destination CTA)
                const int ls = i / NC, dst = i - ls * NC;
                const float e = __fdiv_rn((float)((long long)Z * m.N - 2ll * cnt_sm[2 * ls]), (float)m.N);
                *cluster.map_shared_rank(&en_all[par * S + sys0 + ls], dst) = e;
                if (record) *cluster.map_shared_rank(&mag_all[par * S + sys0 + ls], dst) = m.N - 2ll * cnt_sm[2 * ls + 1];
            }
            }
            last_e = par;
            if (record) last_m = par;
            PROWS_CLK(1);
            // A: all S energies (+ magnetisation sums) of this sweep are in every CTA, and so are the word snapshots; the next
            // exchange is prepared between the arrive and the wait
            asm volatile("barrier.cluster.arrive.release;" ::: "memory");
            cur_edge = pt_edge; cur_logu = pt_logu;
            if (pt_prep) prepare_exchange(pt_event + (pt_this ? 1u : 0u));
            asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
            prepared_next = true;
            PROWS_CLK(2);
        }
        m.energies = en_all + par * S;
        m.mags = mag_all + par * S;
        if (record) {
            if (P > 0) {  // overlap.rs:259-281 with this sweep's pre-exchange system_ids: the pairs of temperature t are handled by the CTA
                          // that folds t (t mod NC), one warp per pair: it fetches both replicas' word snapshots through DSMEM (two
                          // loads per lane in flight together), keeps the XOR words in its own scratch, takes the link terms from
                          // there and leaves the dots in its own shared memory — no second cluster barrier
                uint32_t *X = xs_sm + (size_t)warp * sysw;
                const int n_mine = ((T - rank + NC - 1) / NC) * P;  // (temperature, pair) items of this CTA
                for (int j = warp; j < n_mine; j += n_warps) {
                    const int t = rank + (j / P) * NC, pr = j % P, idx = pr * T + t;
                    const int sa = sid_sm[(2 * pr) * T + t], sb = sid_sm[(2 * pr + 1) * T + t];
                    const int oa = sa / SPC, ob = sb / SPC;
                    const uint32_t *snap_par = snap_sm + (size_t)par * SPC * sysw;
                    const uint32_t *wa = cluster.map_shared_rank(snap_par, oa) + (size_t)(sa - oa * SPC) * sysw;
                    const uint32_t *wb = cluster.map_shared_rank(snap_par, ob) + (size_t)(sb - ob * SPC) * sysw;
                    for (int i = lane; i < sysw; i += 32) X[i] = wa[i] ^ wb[i];
                    __syncwarp();
                    int neg_q = 0, neg_l = 0;
                    if (warp_is_system) {  // lane = word: the forward-neighbour offsets of the lane's own work item apply (no table reads:
                                           // the cluster barrier in front of this phase invalidated L1)
                        const uint32_t x = X[lane];
                        neg_q = __popc(x);
#pragma unroll
                        for (int k = 0; k < Z; k++) {
                            const uint32_t c0 = X[h_n0[0][2 * k]];
                            uint32_t nw = c0;
                            if (dls[k] > 0) nw = (c0 >> 1) | (X[h_n1[0][2 * k]] << 31);
                            else if (dls[k] < 0) nw = (c0 << 1) | (X[h_n1[0][2 * k]] >> 31);
                            neg_l += __popc(x ^ nw);
                        }
                    } else {
                        for (int i = lane; i < sysw; i += 32) {
                            const int r = i / W, w = i - r * W;
                            const uint32_t x = X[i];
                            neg_q += __popc(x);
#pragma unroll
                            for (int k = 0; k < Z; k++) neg_l += __popc(x ^ nbr_word(X, vg.nbr_row[((size_t)r * Z + k) * 2], w, dls[k]));
                        }
                    }
                    neg_q = __reduce_add_sync(0xFFFFFFFFu, neg_q);
                    neg_l = __reduce_add_sync(0xFFFFFFFFu, neg_l);
                    if (lane == 0) {
                        dsp_all[par * PT_ + idx] = m.N - 2ll * neg_q;
                        dlk_all[par * PT_ + idx] = (long long)Z * m.N - 2ll * neg_l;
                    }
                    __syncwarp();
                }
                last_d = par;
                PROWS_CLK(3);
                __syncthreads();  // the dots of this CTA's temperatures are in its shared memory
                PROWS_CLK(4);
            }
            st.dot_spin = dsp_all + par * PT_;
            st.dot_link = dlk_all + par * PT_;
            const int n_own = (T - rank + NC - 1) / NC;  // temperatures rank, rank + NC, ...
            for (int idx = tid; idx < 12 * n_own; idx += NT) {  // mod.rs:543-578: one thread per (temperature, running sum)
                const int t = rank + (idx / 12) * NC;
                fold_spread<true>(
                    m, st, 0, t, idx % 12,
                    [&](int r) { return m.mags[sid_sm[r * T + t]]; },
                    [&](int r) { return m.energies[sid_sm[r * T + t]]; },
                    [&](int p) { return st.dot_spin[p * T + t]; },
                    [&](int p) { return st.dot_link[p * T + t]; });
            }
            __syncthreads();  // the fold read system_ids: the exchange below changes them
            PROWS_CLK(5);
        }
        if (pt_this) {  // mod.rs:748-796, replayed identically by every CTA on its own copy
            if (T >= 2) {
                if (a.pt_schedule == 0) {
                    if (tid < m.R && pt_attempt_edge_pre(m, pt, 0, tid, cur_edge, cur_logu, dbeta_sm[cur_edge])) {
                        // the two systems changed slots (different replicas never touch the same systems)
                        const int s_lo = sid_sm[tid * T + cur_edge] - sys0, s_hi = sid_sm[tid * T + cur_edge + 1] - sys0;
                        if (s_lo >= 0 && s_lo < SPC) slot_of[s_lo] = tid * T + cur_edge;
                        if (s_hi >= 0 && s_hi < SPC) slot_of[s_hi] = tid * T + cur_edge + 1;
                    }
                    __syncthreads();
                } else {
                    if (tid < m.R) pt_exchange_body(m, pt, 0, tid, a.pt_schedule, parity, pt_event);
                    parity = 1 - parity;
                    __syncthreads();
                    rebuild_slots();
                    __syncthreads();
                }
            }
            pt_event++;
            if (!prepared_next && pt_prep) prepare_exchange(pt_event);
        }
        PROWS_CLK(6);
    }
    cluster.sync();  // no CTA leaves (or rewrites global state) while another may still address its shared memory
    const int n_own = (T - rank + NC - 1) / NC;
    for (int i = tid; i < 11 * n_own; i += NT) {
        const int t = rank + (i / 11) * NC, k = i % 11;
        stg.sums[dg * 11 * T + (int64_t)k * T + t] = sums_sm[k * T + t];
    }
    if (last_d >= 0)
        for (int i = tid; i < P * n_own; i += NT) {
            const int t = rank + (i / P) * NC, pr = i % P;
            const_cast<long long *>(stg.dot_spin)[dg * PT_ + pr * T + t] = dsp_all[last_d * PT_ + pr * T + t];
            const_cast<long long *>(stg.dot_link)[dg * PT_ + pr * T + t] = dlk_all[last_d * PT_ + pr * T + t];
        }
    for (int i = tid; i < SPC; i += NT) {
        if (last_e >= 0) mg.energies[dg * S + sys0 + i] = en_all[last_e * S + sys0 + i];
        if (last_m >= 0) mg.mags[dg * S + sys0 + i] = mag_all[last_m * S + sys0 + i];
    }
    if (rank == 0) {
        for (int i = tid; i < S; i += NT) {
            mg.system_ids[dg * S + i] = sid_sm[i];
            ptg.round_trips[dg * S + i] = rt_sm[i];
            ptg.trip_state[dg * S + i] = trip_sm[i];
        }
        for (int i = tid; i < T - 1; i += NT) {
            ptg.edge_attempts[dg * (T - 1) + i] = ea_sm[i];
            ptg.edge_acceptances[dg * (T - 1) + i] = eacc_sm[i];
        }
    }
    for (int i = tid; i < SPC * sysw; i += NT) {  // unpack: +1 = 0x01, -1 = 0xFF
        const uint32_t word = pw[i];
        uint4 *dst = reinterpret_cast<uint4 *>(g_spins + ((int64_t)sys0 * sysw + i) * 32);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t part[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t nib = (word >> (16 * h + 4 * j)) & 15u;
                const uint32_t spread = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                part[j] = 0x01010101u ^ (spread * 0xFEu);
            }
            dst[h] = make_uint4(part[0], part[1], part[2], part[3]);
        }
    }
}

// shared memory of prows_resident_kernel
inline size_t prows_resident_smem(int S, int T, int P, int z, int64_t N) {
    const size_t thr_words = ((size_t)T * (2 * z + 1) + 3) & ~size_t(3);
    const size_t cnt = 2 * (size_t)std::max(S, P * T);
    return thr_words * 4 + resident_scalar_bytes(S, T, P) + (size_t)S * (size_t)(N / 32) * 4 + cnt * 4 + (size_t)T * 4 + 16;
}

}  // namespace pp
