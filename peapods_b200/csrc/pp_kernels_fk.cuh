// pp_kernels_fk.cuh — Fortuin-Kasteleyn cluster update (Swendsen-Wang / Wolff) for the int8 layouts.
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
                    o = philox4x32_10(cached, sweep_index, sys, TAG_FK_BOND, k0, k1);
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
        const u32x4 o = philox4x32_10(0xFFFFFFFFu, sweep_index, sys, TAG_FK_FLIP, k0, k1);
        seed_root = lab[(uint32_t)(((uint64_t)o.y * (uint64_t)N) >> 32)];
    }
    for (int64_t i = tid; i < N; i += FK_THREADS) {
        const uint32_t root = lab[i];
        bool flip;
        if (wolff) {
            flip = root == seed_root;
        } else {
            const u32x4 o = philox4x32_10(root >> 2, sweep_index, sys, TAG_FK_FLIP, k0, k1);
            flip = (pick(o, root & 3u) >> 8) < (1u << 23);
        }
        if (flip) s[i] = (int8_t)-s[i];
    }
}
#endif  // __CUDACC__

}  // namespace pp
