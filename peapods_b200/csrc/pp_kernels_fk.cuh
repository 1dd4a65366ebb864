// pp_kernels_fk.cuh — cluster moves for the int8 layouts: Fortuin-Kasteleyn update (Swendsen-Wang / Wolff) and the Houdayer
// isoenergetic overlap move (second half of the file).
//
// Replaces  clusters/fk.rs:28-171 (fk_update, union-find path; called from simulation/mod.rs:434-470 after the sweep and before
// the measurements).  Bonds between neighbours with s_i s_j J > 0 are activated with probability 1 - exp(-2 / T) (unit
// couplings: an integer cut-off on the 24-bit draw, built on the host with libm like the acceptance tables), clusters are the
// connected components, Swendsen-Wang flips each cluster with probability 1/2, Wolff flips the cluster of one drawn site.
//
// RNG-SPEC (cluster domains; the CPU checker in tests/ consumes the same function): key = the realization's seed; bond (i, d) draws
// from counter {(i z' + d) >> 2, sweep index, system id, TAG_FK_BOND}, lane (i z' + d) & 3; a cluster is named by its smallest
// site r and flips iff the draw of {r >> 2, sweep index, system id, TAG_FK_FLIP}, lane r & 3, is below 2^23; the Wolff seed is
// (out[1] * N) >> 32 of {0xFFFFFFFF, sweep index, system id, TAG_FK_FLIP}.  Naming clusters by their smallest site makes the
// result independent of how the components are found: union-find on the CPU, label propagation here.
#pragma once
#include "pp_device.cuh"

namespace pp {

constexpr uint32_t TAG_FK_BOND = 0x00050000u;
constexpr uint32_t TAG_FK_FLIP = 0x00060000u;
constexpr int FK_THREADS = 512;

#if defined(__CUDACC__)
// One CTA per (realization d, slot k).  labels / bond masks live in shared memory when they fit (smem_sites >= N), else in the
// global scratch `g_lab` ([D * S][N] u32) / `g_bm` ([D * S][N] u8).
__global__ void __launch_bounds__(FK_THREADS)
fk_cluster_kernel(ModelView m, const uint32_t *bond_count /* [T] */, uint32_t sweep_index, int wolff, int64_t smem_sites,
                  uint32_t *g_lab, uint8_t *g_bm) {
    extern __shared__ __align__(16) uint32_t fk_sm[];
    const int64_t N = m.N;
    const int z = m.z, tid = threadIdx.x;
    const int64_t d = blockIdx.x / m.S;
    const int slot = (int)(blockIdx.x % m.S);
    const uint32_t sys = (uint32_t)m.system_ids[d * m.S + slot];
    int8_t *s = m.spins + (d * m.S + sys) * N;
    const bool in_smem = smem_sites >= N;
    uint32_t *lab = in_smem ? fk_sm : g_lab + (int64_t)blockIdx.x * N;
    uint8_t *bm = in_smem ? reinterpret_cast<uint8_t *>(fk_sm + N) : g_bm + (int64_t)blockIdx.x * N;
    const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint32_t count = bond_count[slot % m.T];
    // ---- bonds (fk.rs:107-115: forward bonds, each once)
    for (int64_t i = tid; i < N; i += FK_THREADS) {
        const int si = s[i];
        uint32_t mask = 0u, cached = 0xFFFFFFFFu;
        u32x4 o = {0u, 0u, 0u, 0u};
        for (int dd = 0; dd < z; dd++) {
            const uint32_t j = m.nbr[(size_t)i * 2 * z + 2 * dd];
            int inter = si * (int)s[j];
            if (m.coupling_class == COUP_UNIT) inter *= (int)m.J8[((size_t)d * N + i) * z + dd];
            const uint32_t bond = (uint32_t)(i * z + dd);
            if (inter > 0) {
                if ((bond >> 2) != cached) {
                    cached = bond >> 2;
                    o = philox4x32(cached, sweep_index, sys, TAG_FK_BOND, k0, k1);
                }
                if ((pick(o, bond & 3u) >> 8) < count) mask |= 1u << dd;
            }
        }
        bm[i] = (uint8_t)mask;
        lab[i] = (uint32_t)i;
    }
    __syncthreads();
    // ---- connected components: minimum-label propagation over the active bonds with pointer jumping
    for (;;) {
        int changed = 0;
        for (int64_t i = tid; i < N; i += FK_THREADS) {
            const uint32_t old = lab[i];
            uint32_t best = old;
            const uint32_t mine = bm[i];
            for (int dd = 0; dd < z; dd++) {
                const uint32_t jf = m.nbr[(size_t)i * 2 * z + 2 * dd], jb = m.nbr[(size_t)i * 2 * z + 2 * dd + 1];
                if ((mine >> dd) & 1u) best = min(best, lab[jf]);
                if ((bm[jb] >> dd) & 1u) best = min(best, lab[jb]);
            }
            best = min(best, lab[best]);
            best = min(best, lab[best]);
            if (best < old) {
                atomicMin(&lab[i], best);
                atomicMin(&lab[old], best);  // hook the old root too: whole trees move at once
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    // ---- flips (fk.rs:151-170)
    uint32_t seed_root = 0u;
    if (wolff) {
        const u32x4 o = philox4x32(0xFFFFFFFFu, sweep_index, sys, TAG_FK_FLIP, k0, k1);
        seed_root = lab[(uint32_t)(((uint64_t)o.y * (uint64_t)N) >> 32)];
    }
    for (int64_t i = tid; i < N; i += FK_THREADS) {
        const uint32_t root = lab[i];
        bool flip;
        if (wolff) {
            flip = root == seed_root;
        } else {
            const u32x4 o = philox4x32(root >> 2, sweep_index, sys, TAG_FK_FLIP, k0, k1);
            flip = (pick(o, root & 3u) >> 8) < (1u << 23);
        }
        if (flip) s[i] = (int8_t)-s[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Houdayer isoenergetic cluster move, group size 2 (clusters/overlap.rs:34-56, 146-339; called from simulation/mod.rs:596-746
// after the measurements, energies refreshed for PT only).  One CTA per (realization d, slot t, pair g): the R systems at the
// slot are shuffled (Fisher-Yates on TAG_OC_PAIR draws) and paired; active sites are those where the two replicas differ, bonds
// join active neighbours (no draws), clusters are named by their smallest site.  wolff: the active site with the smallest
// (score, index) -- scores are TAG_OC_SEED draws, so the choice is uniform over the active sites -- is the seed and its cluster
// flips in both replicas; else every cluster of more than one site flips iff its TAG_OC_FLIP draw is below 2^23.
constexpr uint32_t TAG_OC_PAIR = 0x00070000u;
constexpr uint32_t TAG_OC_SEED = 0x00080000u;
constexpr uint32_t TAG_OC_FLIP = 0x00090000u;

__global__ void __launch_bounds__(FK_THREADS)
houdayer_kernel(ModelView m, uint32_t sweep_index, int wolff, int64_t smem_sites, uint32_t *g_lab, uint8_t *g_bm) {
    extern __shared__ __align__(16) uint32_t fk_sm[];
    __shared__ unsigned long long best_sh;
    __shared__ int sys_sh[2];
    const int64_t N = m.N;
    const int z = m.z, tid = threadIdx.x;
    const int g = (int)(blockIdx.x % m.P);
    const int t = (int)((blockIdx.x / m.P) % m.T);
    const int64_t d = blockIdx.x / ((int64_t)m.P * m.T);
    const bool in_smem = smem_sites >= N;
    uint32_t *lab = in_smem ? fk_sm : g_lab + (int64_t)blockIdx.x * N;
    uint8_t *act = in_smem ? reinterpret_cast<uint8_t *>(fk_sm + N) : g_bm + (int64_t)blockIdx.x * N;  // bit 0: active, bit 1: has an active neighbour
    const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint32_t stream = (uint32_t)(t * m.P + g);
    if (tid == 0) {  // overlap.rs:45-49: the slot's systems in replica order, shuffled
        int sys[64];
        for (int k = 0; k < m.R; k++) sys[k] = m.system_ids[d * m.S + k * m.T + t];
        for (int i = m.R - 1; i >= 1; i--) {
            const u32x4 o = philox4x32((uint32_t)i, sweep_index, (uint32_t)t, TAG_OC_PAIR, k0, k1);
            const int j = (int)(((uint64_t)o.x * (uint64_t)(i + 1)) >> 32);
            const int tmp = sys[i]; sys[i] = sys[j]; sys[j] = tmp;
        }
        sys_sh[0] = sys[2 * g];
        sys_sh[1] = sys[2 * g + 1];
        best_sh = ~0ull;
    }
    __syncthreads();
    int8_t *a = m.spins + (d * m.S + sys_sh[0]) * N, *b = m.spins + (d * m.S + sys_sh[1]) * N;
    for (int64_t i = tid; i < N; i += FK_THREADS) {
        act[i] = a[i] != b[i] ? 1 : 0;
        lab[i] = (uint32_t)i;
    }
    __syncthreads();
    for (int64_t i = tid; i < N; i += FK_THREADS) {  // sites with an active neighbour (clusters of more than one site)
        if (!(act[i] & 1)) continue;
        bool multi = false;
        for (int dd = 0; dd < 2 * z; dd++) {
            const uint32_t j = m.nbr[(size_t)i * 2 * z + dd];
            multi = multi || (j != (uint32_t)i && (act[j] & 1));
        }
        if (multi) act[i] |= 2;
    }
    __syncthreads();
    for (;;) {  // connected components of the active sites
        int changed = 0;
        for (int64_t i = tid; i < N; i += FK_THREADS) {
            if (!(act[i] & 2)) continue;
            const uint32_t old = lab[i];
            uint32_t best = old;
            for (int dd = 0; dd < 2 * z; dd++) {
                const uint32_t j = m.nbr[(size_t)i * 2 * z + dd];
                if (act[j] & 1) best = min(best, lab[j]);
            }
            best = min(best, lab[best]);
            best = min(best, lab[best]);
            if (best < old) {
                atomicMin(&lab[i], best);
                atomicMin(&lab[old], best);
                changed = 1;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (wolff) {  // overlap.rs:245-256
        unsigned long long best = ~0ull;
        for (int64_t i = tid; i < N; i += FK_THREADS) {
            if (!(act[i] & 1)) continue;
            const u32x4 o = philox4x32((uint32_t)i >> 2, sweep_index, stream, TAG_OC_SEED, k0, k1);
            const unsigned long long score = ((unsigned long long)(pick(o, (uint32_t)i & 3u) >> 8) << 32) | (unsigned long long)i;
            best = score < best ? score : best;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            best = other < best ? other : best;
        }
        if ((tid & 31) == 0 && best != ~0ull) atomicMin(&best_sh, best);
        __syncthreads();
        if (best_sh == ~0ull) return;  // no active site
        const uint32_t root = lab[(uint32_t)(best_sh & 0xFFFFFFFFull)];
        for (int64_t i = tid; i < N; i += FK_THREADS)
            if ((act[i] & 1) && lab[i] == root) { a[i] = (int8_t)-a[i]; b[i] = (int8_t)-b[i]; }
    } else {  // overlap.rs:293-307
        for (int64_t i = tid; i < N; i += FK_THREADS) {
            if ((act[i] & 3) != 3) continue;
            const uint32_t root = lab[i];
            const u32x4 o = philox4x32(root >> 2, sweep_index, stream, TAG_OC_FLIP, k0, k1);
            if ((pick(o, root & 3u) >> 8) < (1u << 23)) { a[i] = (int8_t)-a[i]; b[i] = (int8_t)-b[i]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same move for the multispin layout (Wolff mode): one CTA per (word group g, slot t, pair): 32 realizations at once.
// Everything that is lane-uniform is drawn with the GROUP key (as the multispin sweeps do): the pairing of the replica ladders
// at the slot and the per-site seed scores; the seed of lane l is its own active site with the smallest (score, site), found
// by a per-lane scan, and the 32 clusters grow together as a bit-parallel flood fill: C |= X & (C of the 2z' neighbours) until
// nothing changes.  Flipping the cluster in both replicas is an XOR of the two systems' words with C.
// shared memory (u32 words): X[N] active masks | C[N] cluster masks | score[N] (by logical site) | nbr16[N * 2z'] (u16).
#ifndef PP_OC_THREADS
#define PP_OC_THREADS 512
#endif
constexpr int OC_THREADS = PP_OC_THREADS;  // threads of msc_houdayer_kernel (the fill is bound by shared-memory latency: more warps)
constexpr int OC_WARPS = OC_THREADS / 32;

__global__ void nbr_to_u16_kernel(const uint32_t *nbr, uint16_t *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint16_t)nbr[i];
}

// ZT > 0: the number of forward directions is the compile-time constant ZT (the neighbour gathers unroll); 0: any z'.
// nbr16 = the storage-space neighbour table as u16 [N][2z'] (built once per handle; N <= 65536).
template <int ZT>
__global__ void __launch_bounds__(OC_THREADS)
msc_houdayer_kernel(ModelView m, const uint16_t *nbr16, const uint16_t *site16 /* [N] storage index -> site */, uint32_t sweep_index,
                    int64_t group_offset) {
    extern __shared__ __align__(16) uint32_t fk_sm[];
    __shared__ unsigned long long best_sh[OC_WARPS][32];
    __shared__ int pair_sh[2];
    const int64_t N = m.N;
    const int z2 = ZT > 0 ? 2 * ZT : 2 * m.z, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t *X = fk_sm, *Cm = fk_sm + N, *score = fk_sm + 2 * N;
    uint16_t *nb = reinterpret_cast<uint16_t *>(fk_sm + 3 * N);
    const int pg = (int)(blockIdx.x % m.P);
    const int t = (int)((blockIdx.x / m.P) % m.T);
    const int64_t g = blockIdx.x / ((int64_t)m.P * m.T);
    const uint64_t key = msc_group_key(m.seed, (uint64_t)(group_offset + g));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint32_t stream = (uint32_t)(t * m.P + pg);
    if (tid == 0) {  // overlap.rs:45-49: the slot's replica ladders, shuffled (the same pairing for the 32 lanes)
        int idx[64];
        for (int k = 0; k < m.R; k++) idx[k] = k;
        for (int i = m.R - 1; i >= 1; i--) {
            const u32x4 o = philox4x32((uint32_t)i, sweep_index, (uint32_t)t, TAG_OC_PAIR, k0, k1);
            const int j = (int)(((uint64_t)o.x * (uint64_t)(i + 1)) >> 32);
            const int tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp;
        }
        pair_sh[0] = idx[2 * pg];
        pair_sh[1] = idx[2 * pg + 1];
    }
    __syncthreads();
    uint32_t *A = m.words + ((g * m.S + (int64_t)pair_sh[0] * m.T + t) * N);
    uint32_t *B = m.words + ((g * m.S + (int64_t)pair_sh[1] * m.T + t) * N);
    for (int64_t i = tid; i < N; i += OC_THREADS) {
        X[i] = A[i] ^ B[i];
        Cm[i] = 0u;
    }
    {  // the neighbour table: N * 2z' u16 = a multiple of 4 bytes (N is even in every multispin layout)
        const uint32_t *src = reinterpret_cast<const uint32_t *>(nbr16);
        uint32_t *dst = reinterpret_cast<uint32_t *>(nb);
        for (int64_t i = tid; i < N * z2 / 2; i += OC_THREADS) dst[i] = __ldg(src + i);
    }
    for (int64_t q = tid; q < (N + 3) / 4; q += OC_THREADS) {  // scores are drawn per logical site and kept by storage position
        const u32x4 o = philox4x32((uint32_t)q, sweep_index, stream, TAG_OC_SEED, k0, k1);
        for (int j = 0; j < 4; j++)
            if (4 * q + j < N) score[m.perm ? m.perm[4 * q + j] : (uint32_t)(4 * q + j)] = pick(o, (uint32_t)j) >> 8;
    }
    __syncthreads();
    // Seed of lane `lane` = its active site with the smallest (score, site).  The scores are lane-uniform, so the few hundred
    // lowest-score positions are collected once and every lane looks for its first active one among them; a lane that has no
    // active site there (few active sites at all) falls back to scanning every word.
    __shared__ uint32_t cand_n;
    __shared__ uint16_t cand[1024];
    if (tid == 0) cand_n = 0u;
    __syncthreads();
    const uint32_t cut = (uint32_t)min((int64_t)(1 << 24), ((int64_t)256 << 24) / N);  // expect ~256 candidates
    for (int64_t p = tid; p < N; p += OC_THREADS)
        if (score[p] < cut) {
            const uint32_t k = atomicAdd(&cand_n, 1u);
            if (k < 1024u) cand[k] = (uint16_t)p;
        }
    __syncthreads();
    {
        unsigned long long best = ~0ull;
        const uint32_t n_cand = cand_n;
        if (n_cand <= 1024u)
            for (uint32_t k = w; k < n_cand; k += OC_WARPS) {
                const uint32_t p = cand[k];
                if ((X[p] >> lane) & 1u) {
                    const unsigned long long sc = ((unsigned long long)score[p] << 32) | (unsigned long long)__ldg(site16 + p);
                    best = sc < best ? sc : best;
                }
            }
        best_sh[w][lane] = best;
    }
    __syncthreads();
    {  // fallback: lanes without a candidate (warp-uniform test, so whole warps skip the scan)
        unsigned long long have = ~0ull;
        for (int k = 0; k < OC_WARPS; k++) have = best_sh[k][lane] < have ? best_sh[k][lane] : have;
        const bool need = have == ~0ull;
        __syncthreads();  // every warp has read the candidate minima before anyone overwrites its slot
        if (__any_sync(0xFFFFFFFFu, need)) {
            unsigned long long best = ~0ull;
            const int64_t per = (N + OC_WARPS - 1) / OC_WARPS, p0 = w * per, p1 = min(N, p0 + per);
            for (int64_t p = p0; p < p1; p++) {
                if (need && ((X[p] >> lane) & 1u)) {
                    const unsigned long long sc = ((unsigned long long)score[p] << 32) | (unsigned long long)__ldg(site16 + p);
                    best = sc < best ? sc : best;
                }
            }
            if (need) best_sh[w][lane] = best;
        }
    }
    __syncthreads();
    if (w == 0) {
        unsigned long long best = best_sh[0][lane];
        for (int k = 1; k < OC_WARPS; k++) best = best_sh[k][lane] < best ? best_sh[k][lane] : best;
        if (best != ~0ull) {
            const uint32_t site = (uint32_t)(best & 0xFFFFFFFFull);
            atomicOr(&Cm[m.perm ? m.perm[site] : site], 1u << lane);
        }
    }
    __syncthreads();
    for (;;) {  // bit-parallel flood fill over the active sites (overlap.rs:316-327: the seed's connected component)
        int changed = 0;
        for (int64_t p = tid; p < N; p += OC_THREADS) {
            const uint32_t x = X[p], c = Cm[p];
            if (x & ~c) {
                uint32_t n = 0u;
#pragma unroll
                for (int k = 0; k < (ZT > 0 ? 2 * ZT : 16); k++)
                    if (k < z2) n |= Cm[nb[p * z2 + k]];
                const uint32_t c2 = c | (n & x);
                if (c2 != c) {
                    Cm[p] = c2;
                    changed = 1;
                }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int64_t p = tid; p < N; p += OC_THREADS) {
        const uint32_t c = Cm[p];
        if (c) {
            A[p] ^= c;
            B[p] ^= c;
        }
    }
}
#endif  // __CUDACC__

}  // namespace pp
