// pp_kernels_msc.cuh — multispin-coded layout: one u32 word holds the same site of 32 disorder
// realizations (bit = 1: spin -1), words are stored slot-major: W[group][replica*T + slot][site].
// Couplings are sign words Jw[group][d][site] (bit = 1: J = -1).  All 32 lanes of a word sit at the
// same temperature, so one 24-bit draw and one row of the acceptance table serve the whole word
// (RNG-SPEC: TAG_SWEEP_MSC, stream = replica*T + slot, key = word-group key).
//
// Arithmetic restated bit-sliced from spin-sim/src/mcmc/sweep.rs:8-19 and :170-185:
//   unsat_l = #{neighbours j : s_i s_j J_ij = -1}   (vertical counter over the 2z' bond words)
//   ec_l    = -s_i h = 2*unsat_l - 2z'              -> table index ec + 2z' = 2*unsat_l
//   flip_l  = draw < table[t][2*unsat_l]
// Parallel tempering swaps labels in the reference (mcmc/tempering.rs:93); here the labels are
// swapped too (system_ids) and the affected lanes of the two slot words are exchanged
// (x = (A^B)&mask; A^=x; B^=x), so that every word keeps a single temperature.
#pragma once
#include "pp_device.cuh"

namespace pp {

constexpr int MSC_BLOCK = 256;
constexpr int MSC_VC_PLANES = 20;  // per-thread vertical counter capacity 2^20 - 1 adds (pp_create checks ceil(N / 256) * z against it)

// ---- vertical (bit-sliced) counters ----------------------------------------------------------
template <int K>
struct VCount {
    uint32_t p[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int b = 0; b < K; b++) p[b] = 0;
    }
    __device__ __forceinline__ void add(uint32_t x) {
        uint32_t c = x;
#pragma unroll
        for (int b = 0; b < K; b++) {
            uint32_t t = p[b] & c;
            p[b] ^= c;
            c = t;
            if (c == 0) break;
        }
    }
};

// Sum the K-plane counters of all threads of the block; thread l < 32 returns the total of lane l.
// scratch: [nwarps][K+5] words.
template <int K>
__device__ __forceinline__ uint32_t block_lane_totals(const VCount<K> &vc, uint32_t *scratch) {
    constexpr int KW = K + 5;
    uint32_t a[KW];
#pragma unroll
    for (int b = 0; b < KW; b++) a[b] = b < K ? vc.p[b] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t carry = 0;
#pragma unroll
        for (int b = 0; b < KW; b++) {
            uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, a[b], o);
            uint32_t s = a[b] ^ y ^ carry;
            carry = (a[b] & y) | (carry & (a[b] ^ y));
            a[b] = s;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int b = 0; b < KW; b++) scratch[wid * KW + b] = a[b];
    __syncthreads();
    uint32_t total = 0;
    if (wid == 0)
        for (int w = 0; w < nw; w++)
#pragma unroll
            for (int b = 0; b < KW; b++) total += ((scratch[w * KW + b] >> lane) & 1u) << b;
    return total;
}

// unsat-count planes (c0..c3) of 2z bond words
__device__ __forceinline__ void count_bonds(const uint32_t *bond, int n, uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                            uint32_t &c3) {
    c0 = c1 = c2 = c3 = 0;
    if (n == 6) {
        uint32_t s1 = bond[0] ^ bond[1] ^ bond[2], k1 = (bond[0] & bond[1]) | (bond[2] & (bond[0] ^ bond[1]));
        uint32_t s2 = bond[3] ^ bond[4] ^ bond[5], k2 = (bond[3] & bond[4]) | (bond[5] & (bond[3] ^ bond[4]));
        c0 = s1 ^ s2;
        uint32_t k3 = s1 & s2;
        c1 = k1 ^ k2 ^ k3;
        c2 = (k1 & k2) | (k3 & (k1 ^ k2));
    } else if (n == 4) {
        uint32_t s1 = bond[0] ^ bond[1] ^ bond[2], k1 = (bond[0] & bond[1]) | (bond[2] & (bond[0] ^ bond[1]));
        c0 = s1 ^ bond[3];
        uint32_t k3 = s1 & bond[3];
        c1 = k1 ^ k3;
        c2 = k1 & k3;
    } else {
        for (int k = 0; k < n; k++) {
            uint32_t c = bond[k], t;
            t = c0 & c; c0 ^= c; c = t;
            t = c1 & c; c1 ^= c; c = t;
            t = c2 & c; c2 ^= c; c = t;
            c3 ^= c;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// init: word (g, pos, i), lane l = realization 32g+l: bit = INIT draw of (key(l), system = pos, site i) < 2^23
// grid.x = G*S, grid.y = quads of sites
__global__ void msc_init_kernel(ModelView m) {
    const int64_t wsys = blockIdx.x;  // g*S + pos
    const int64_t g = wsys / m.S;
    const uint32_t pos = (uint32_t)(wsys % m.S);
    const int64_t q = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (q * 4 >= m.N) return;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int l = 0; l < 32; l++) {
        const int64_t d = g * 32 + l;
        if (d >= m.D) break;  // padding lanes stay +1
        const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
        const u32x4 o = philox4x32((uint32_t)q, 0u, pos, TAG_INIT, (uint32_t)key, (uint32_t)(key >> 32));
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] |= ((pick(o, j) >> 8) < (1u << 23) ? 1u : 0u) << l;
    }
    uint32_t *dst = m.words + wsys * m.N;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (q * 4 + j < m.N) dst[m.perm ? m.perm[q * 4 + j] : (uint32_t)(q * 4 + j)] = w[j];
}

// ---------------------------------------------------------------------------------------------
// K2 (generic lattice): one CTA per word-system; all colour classes of one sweep inside the kernel,
// the system's N words staged in shared memory when they fit.  Optional epilogue: per-lane
// energy (forward unsatisfied bonds) and magnetisation (down spins) from the staged words.
template <bool IN_SMEM>
__global__ void __launch_bounds__(MSC_BLOCK)
msc_sweep_kernel(ModelView m, uint32_t sweep_index, int n_sweeps, int want_energy, int want_mags, int64_t group_offset) {
    extern __shared__ uint32_t smem[];
    __shared__ uint32_t scratch[(MSC_BLOCK / 32) * (MSC_VC_PLANES + 5)];
    __shared__ uint32_t thr[40];  // acceptance count for unsat = u: lut[t][2u]
    const int64_t wsys = blockIdx.x;  // g*S + pos
    const int64_t g = wsys / m.S;
    const uint32_t pos = (uint32_t)(wsys % m.S);
    const int t = pos % m.T;
    const int z = m.z, z2 = 2 * m.z;
    const int64_t N = m.N;
    uint32_t *gw = m.words + wsys * N;
    const uint32_t *J = m.Jw ? m.Jw + g * z * N : nullptr;
    uint32_t *w = IN_SMEM ? smem : gw;
    if (IN_SMEM) {
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) w[i] = gw[i];
    }
    if (threadIdx.x <= z2) thr[threadIdx.x] = m.lut[t * (4 * z + 1) + 2 * threadIdx.x];
    __syncthreads();
    const uint64_t key = msc_group_key(m.seed, (uint64_t)(group_offset + g));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);

    for (int sw = 0; sw < n_sweeps; sw++) {
        for (int c = 0; c < m.n_colours; c++) {
            const uint32_t cs = m.colour_start[c], ce = m.colour_start[c + 1];
            const uint32_t nq = (ce - cs + 3) >> 2;
            for (uint32_t q = threadIdx.x; q < nq; q += blockDim.x) {
                const u32x4 o = philox4x32(q, sweep_index + (uint32_t)sw, pos, TAG_SWEEP_MSC | (uint32_t)c, k0, k1);
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const uint32_t p = cs + q * 4 + l;
                    if (p >= ce) break;
                    const uint32_t i = m.order[p];
                    const uint32_t draw = pick(o, l) >> 8;
                    const uint32_t *nb = m.nbr + (size_t)i * z2;
                    const uint32_t wi = w[i];
                    uint32_t bond[32];
                    for (int k = 0; k < z; k++) {
                        const uint32_t jf = nb[2 * k], jb = nb[2 * k + 1];
                        bond[2 * k] = wi ^ w[jf] ^ (J ? J[(size_t)k * N + i] : 0u);
                        bond[2 * k + 1] = wi ^ w[jb] ^ (J ? J[(size_t)k * N + jb] : 0u);
                    }
                    uint32_t c0, c1, c2, c3;
                    count_bonds(bond, z2, c0, c1, c2, c3);
                    uint32_t flip = 0;
                    for (int u = 0; u <= z2; u++) {
                        if (draw < thr[u]) {
                            uint32_t eq = ((u & 1) ? c0 : ~c0) & ((u & 2) ? c1 : ~c1) & ((u & 4) ? c2 : ~c2) &
                                          ((u & 8) ? c3 : ~c3);
                            flip |= eq;
                        }
                    }
                    w[i] = wi ^ flip;
                }
            }
            __syncthreads();
        }
    }

    if (want_energy || want_mags) {
        VCount<MSC_VC_PLANES> vu, vd;
        vu.clear();
        vd.clear();
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            const uint32_t wi = w[i];
            if (want_mags) vd.add(wi);
            if (want_energy) {
                const uint32_t *nb = m.nbr + (size_t)i * z2;
                for (int k = 0; k < z; k++) vu.add(wi ^ w[nb[2 * k]] ^ (J ? J[(size_t)k * N + i] : 0u));
            }
        }
        const uint32_t lane = threadIdx.x;
        const int64_t d = g * 32 + lane;
        if (want_energy) {
            uint32_t unsat = block_lane_totals(vu, scratch);
            if (threadIdx.x < 32 && d < m.D) {
                const int sys = m.system_ids[d * m.S + pos];
                const long long e_int = (long long)N * z - 2ll * unsat;  // sum_i sum_d s s J  (energy.rs:103-107)
                m.energies[d * m.S + sys] = __fdiv_rn((float)e_int, (float)N);
            }
        }
        if (want_mags) {
            uint32_t down = block_lane_totals(vd, scratch);
            if (threadIdx.x < 32 && d < m.D) {
                const int sys = m.system_ids[d * m.S + pos];
                m.mags[d * m.S + sys] = (long long)N - 2ll * down;
            }
        }
    }
    if (IN_SMEM) {
        __syncthreads();
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) gw[i] = w[i];
    }
}

// ---------------------------------------------------------------------------------------------
// K6 (MSC): per (group, pair, slot): x = A ^ B (bit = 1: q_i = -1)
//   dot_spin_l = N   - 2 * #{i : x_i}                      (overlap.rs:266-271)
//   dot_link_l = N z - 2 * #{(i,d) : x_i ^ x_fwd(i,d)}     (overlap.rs:272-276)
__global__ void __launch_bounds__(MSC_BLOCK)
msc_overlap_kernel(ModelView m, long long *dot_spin, long long *dot_link) {
    __shared__ uint32_t scratch[(MSC_BLOCK / 32) * (MSC_VC_PLANES + 5)];
    const int64_t idx = blockIdx.x;  // (g*P + p)*T + t
    const int t = (int)(idx % m.T);
    const int p = (int)((idx / m.T) % m.P);
    const int64_t g = idx / ((int64_t)m.T * m.P);
    const int64_t N = m.N;
    const int z = m.z, z2 = 2 * m.z;
    const uint32_t *a = m.words + (g * m.S + (2 * p) * m.T + t) * N;
    const uint32_t *b = m.words + (g * m.S + (2 * p + 1) * m.T + t) * N;
    VCount<MSC_VC_PLANES> vs, vl;
    vs.clear();
    vl.clear();
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const uint32_t x = a[i] ^ b[i];
        vs.add(x);
        const uint32_t *nb = m.nbr + (size_t)i * z2;
        for (int k = 0; k < z; k++) {
            const uint32_t n = nb[2 * k];
            vl.add(x ^ a[n] ^ b[n]);
        }
    }
    const uint32_t cs = block_lane_totals(vs, scratch);
    const uint32_t cl = block_lane_totals(vl, scratch);
    const int64_t d = g * 32 + threadIdx.x;
    if (threadIdx.x < 32 && d < m.D) {
        const int64_t o = (d * m.P + p) * m.T + t;
        dot_spin[o] = (long long)N - 2ll * cs;
        dot_link[o] = (long long)N * z - 2ll * cl;
    }
}

// ---------------------------------------------------------------------------------------------
// Exchange the lanes whose PT swap was accepted between the two slot words of every edge, in the
// order the schedule attempted the edges.  grid.x = G*R ladders, grid.y = site chunks.
__global__ void msc_apply_swaps_kernel(ModelView m, const uint32_t *swap_mask, int schedule, int first_parity) {
    const int64_t lad = blockIdx.x;  // g*R + r
    const int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (i >= m.N) return;
    const uint32_t *mask = swap_mask + lad * (m.T - 1);
    uint32_t *base = m.words + lad * m.T * m.N + i;  // (g*S + r*T + slot)*N + i
    const int n_pass = schedule == 0 ? 1 : 2;
    for (int pi = 0; pi < n_pass; pi++) {
        const int start = schedule == 0 ? 0 : (pi == 0 ? first_parity : 1 - first_parity);
        const int step = schedule == 0 ? 1 : 2;
        for (int e = start; e < m.T - 1; e += step) {
            const uint32_t mk = mask[e];
            if (mk == 0) continue;
            uint32_t a = base[(int64_t)e * m.N], b = base[(int64_t)(e + 1) * m.N];
            const uint32_t x = (a ^ b) & mk;
            base[(int64_t)e * m.N] = a ^ x;
            base[(int64_t)(e + 1) * m.N] = b ^ x;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pack / unpack between the int8 system-major view of one realization and its lane of the words
__global__ void msc_unpack_kernel(ModelView m, int64_t d, int8_t *out /* [S][N] system-major */) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)m.S * m.N) return;
    const int pos = (int)(gid / m.N);
    const int64_t i = gid % m.N;
    const int sys = m.system_ids[d * m.S + pos];
    const uint32_t w = m.words[((d >> 5) * m.S + pos) * m.N + (m.perm ? m.perm[i] : (uint32_t)i)];
    out[(int64_t)sys * m.N + i] = ((w >> (d & 31)) & 1u) ? (int8_t)-1 : (int8_t)1;
}

__global__ void msc_pack_kernel(ModelView m, int64_t d, const int8_t *in /* [S][N] system-major */) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)m.S * m.N) return;
    const int pos = (int)(gid / m.N);
    const int64_t i = gid % m.N;
    const int sys = m.system_ids[d * m.S + pos];
    uint32_t *w = &m.words[((d >> 5) * m.S + pos) * m.N + (m.perm ? m.perm[i] : (uint32_t)i)];
    const uint32_t bit = 1u << (d & 31);
    *w = in[(int64_t)sys * m.N + i] < 0 ? (*w | bit) : (*w & ~bit);
}

// couplings: float [D][N][z] -> sign words [G][z][N]; one thread per word
__global__ void msc_pack_couplings_kernel(const float *Jf, uint32_t *Jw, int64_t D, int64_t N, int z, const uint32_t *perm) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t G = (D + 31) / 32;
    if (gid >= G * z * N) return;
    const int64_t i = gid % N;
    const int k = (int)((gid / N) % z);
    const int64_t g = gid / (N * z);
    uint32_t w = 0;
    for (int l = 0; l < 32; l++) {
        const int64_t d = g * 32 + l;
        if (d < D && Jf[(d * N + i) * z + k] < 0.0f) w |= 1u << l;
    }
    Jw[(g * z + k) * N + (perm ? perm[i] : (uint32_t)i)] = w;
}

}  // namespace pp
