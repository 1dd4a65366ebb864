// pp_kernels_swords.cuh — fp32 couplings with one BIT per spin: the same site of 32 systems of a realization in one word.
//
// Gaussian (any fp32) couplings differ from realization to realization, so the multispin layout of the +-J path (32 realizations
// per word) does not apply; but the S = R * T systems of ONE realization share its couplings (realization.rs:132-183).  Word
// (d, w, i) of `words[D][SW][N]` (SW = ceil(S / 32), natural site order) holds the spin of the systems 32 w .. 32 w + 31 at site i,
// bit 1 = spin -1.  A thread that owns site i of word slab w loads 1 + 2z' spin words and 2z' couplings ONCE and then walks the 32
// systems with register-only work: per attempt 2z' x (shift, sign-xor, add) for the local field in the reference's order
// (sweep.rs:8-19: direction-major, forward then backward, no fused multiply-add: a product with a +-1 spin is the coupling with its
// sign bit flipped), the log-form rule (sweep.rs:35-48, 247-257 / 279-282) and a quarter of a Philox call.  The int8 row kernel
// (pp_kernels_rows.cuh) issues 100-115 instructions per attempt (byte extraction, per-system addressing, coupling loads per four
// systems); this form issues about a third of that, and reads the couplings once per 32 systems instead of once per four.
//
// Draws: the system-quad mapping of RNG-SPEC (pp_rng.cuh TAG_SWEEP_SYSQ): counter = {colour rank of the site, sweep index,
// system >> 2, tag | colour}, system s takes out[s & 3] >> 8 — one call per four lanes of a word.  Systems stay in their lanes; parallel tempering swaps labels (system_ids, tempering.rs:93) and the kernel
// rebuilds the lane -> temperature map of its word from system_ids at the start of every launch.
//
// Energies (needed by every exchange): the colour-1 pass of a two-colour lattice touches every bond exactly once, so it also adds up
// s_i h_i after the update per lane (energy.rs:99-108), each term rounded to an integer number of 1 / escale (a power of two) on the
// spot, so that every addition is an integer addition and the energies do not depend on threads, blocks or batch size.  Magnetisations and replica overlaps work on a
// TRANSPOSED view produced on recorded sweeps (swords_transpose_kernel: 32 x 32 bit transposes in registers -> `tbits[D][S][N / 32]`,
// one bit per spin, system-major), where a replica pair's q and q_link are XORs and popcounts of whole row words
// (overlap.rs:259-281).  The int8 array stays as a scratch VIEW for get_spins / set_spins / the cluster moves.
// Eligibility (pp_create): fp32 coupling class, two-colour row-alternating lattice (rows_plan), z' = 2 or 3, last extent a multiple
// of 32 or below 32 with N a multiple of 32 (8, 16, 24), offsets that move by at most one site along the rows, at least 16 systems
// per realization.
#pragma once
#include "pp_device.cuh"
#include "pp_kernels_rows.cuh"

namespace pp {

struct SWordsView {
    uint32_t *words;          // [D][SW][N]
    uint32_t *tbits;          // [D][S][N / 32] transposed view (valid after swords_transpose_kernel)
    long long *acc_e;         // [D][SW][32] in-sweep bond sums in units of 1 / escale (kept zero between launches)
    long long *acc_m;         // [D][SW][32] down-spin counts (kept zero between launches)
    unsigned int *arrive_e;   // [D][SW]
    unsigned int *arrive_m;   // [D][SW]
    float4 *lane_t;           // [D][SW][32] per system: (T / 2, ln 2 * T / 2, -24 ln 2 * T / 2, 0) of the temperature it sits at
    int SW;
    float escale;
};

#ifndef PP_SW_THREADS
#define PP_SW_THREADS 128
#endif
#ifndef PP_SW_MINB
#define PP_SW_MINB 8   // resident blocks per SM the plain colour pass is compiled for (registers per thread follow)
#endif
#ifndef PP_SW_MINB_E
#define PP_SW_MINB_E 6 // the same for the pass that also adds up the bond sums (32 more accumulators)
#endif
constexpr int SW_THREADS = PP_SW_THREADS;
constexpr int SW_SPT = 8;  // sites per thread of a colour pass (at most; the host lowers it for small batches so that the grid fills the GPU)

#if defined(__CUDACC__)

__device__ __forceinline__ float sw_lg2(const float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// out[l] bit x = in[x] bit l (32 x 32 bit transpose in registers, five butterfly stages)
__host__ __device__ __forceinline__ void sw_transpose32(uint32_t (&a)[32]) {
    uint32_t mask = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k + j]) & mask;
                a[k] ^= t << j;
                a[k + j] ^= t;
            }
        }
        mask ^= mask << (j >> 1);
    }
}

// One colour pass (UPDATE) and / or the bond sums of the colour's sites (EACC).  grid = (D * SW, ceil(N / 2 / (SW_THREADS * spt))),
// spt <= SW_SPT sites per thread.
template <int Z, bool GIBBS, bool UPDATE, bool EACC, bool EXACT>
__global__ void __launch_bounds__(SW_THREADS, EACC ? PP_SW_MINB_E : PP_SW_MINB)
swords_sweep_kernel(ModelView m, RowsView v, SWordsView sv, int colour, uint32_t sweep_index, int spt) {
    __shared__ float ht_sm[32];
    __shared__ float2 kt_sm[32];  // production-mode threshold of a lane: kt.x * log2(draw) + kt.y = (T / 2) ln u
    __shared__ int blk_sm[32];
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const int64_t d = blockIdx.x / sv.SW;
    const int w = (int)(blockIdx.x % sv.SW);
    const int nl = min(32, m.S - 32 * w);
    if (tid < 32) {  // lane -> temperature of this launch (swords_lane_temps_kernel)
        const float4 lt = UPDATE ? sv.lane_t[((size_t)d * sv.SW + w) * 32 + tid] : make_float4(1.0f, 1.0f, -24.0f, 0.0f);
        ht_sm[tid] = lt.x;
        kt_sm[tid] = make_float2(lt.y, lt.z);
        blk_sm[tid] = 0;
    }
    __syncthreads();
    const int L = v.L, Lh = L >> 1;
    const uint32_t n_act = (uint32_t)v.n_rows * (uint32_t)Lh;  // two colours: every row holds L / 2 sites of each
    uint32_t *W = sv.words + ((size_t)d * sv.SW + w) * m.N;
    const float *J = m.Jf + (size_t)d * m.N * Z;
    const uint64_t key = v.keys[d];
    const PhiloxKeys ks = philox_keys((uint32_t)key, (uint32_t)(key >> 32));
    const uint32_t vmask = nl == 32 ? 0xFFFFFFFFu : ((1u << nl) - 1u);
    const uint32_t tagc = TAG_SWEEP_SYSQ | (uint32_t)colour;
    const float escale = sv.escale;
    // per lane: sum over this thread's sites of round(s h escale) after the update, every term carried by the low mantissa bits of
    // s h * escale + 1.5 * 2^23 (round to nearest; |s h| escale < 2^21): integer sums from the first addition on, so the energies
    // do not depend on how sites are grouped into threads, blocks or batches
    int eacc[EACC ? 32 : 1];
#pragma unroll
    for (int l = 0; l < (EACC ? 32 : 1); l++) eacc[l] = 0;
    uint32_t nterms = 0;
    for (int it = 0; it < spt; it++) {
        const uint32_t c = (blockIdx.y * (uint32_t)spt + (uint32_t)it) * SW_THREADS + tid;  // = the site's rank inside its colour class
        if (c >= n_act) break;
        const uint32_t r = c / (uint32_t)Lh, j = c - r * (uint32_t)Lh;
        const int off = (int)v.row_a[r] == colour ? 0 : 1;
        const int x = 2 * (int)j + off;
        const uint32_t i = r * (uint32_t)L + (uint32_t)x;
        const uint32_t C = W[i];
        uint32_t F[Z], B[Z];
        float Jf[Z], Jb[Z];
#pragma unroll
        for (int kk = 0; kk < Z; kk++) {
            const uint32_t rf = v.nbr_row[((size_t)r * Z + kk) * 2], rb = v.nbr_row[((size_t)r * Z + kk) * 2 + 1];
            const int dl = v.dl[kk];
            int xf = x + dl, xb = x - dl;
            xf -= xf >= L ? L : 0; xf += xf < 0 ? L : 0;
            xb -= xb >= L ? L : 0; xb += xb < 0 ? L : 0;
            const uint32_t jb = rb * (uint32_t)L + (uint32_t)xb;
            F[kk] = W[rf * (uint32_t)L + (uint32_t)xf];
            B[kk] = W[jb];
            Jf[kk] = J[(size_t)i * Z + kk];   // lattice.rs:4-8: bond (i, k) is stored at its lower site
            Jb[kk] = J[(size_t)jb * Z + kk];
        }
#pragma unroll
        for (int kk = 0; kk < Z; kk++) {  // bit l: the neighbour's spin differs from the site's own in system l
            F[kk] ^= C;
            B[kk] ^= C;
        }
        nterms++;
        // Lanes from 31 down to 0: the decision of a lane is the SIGN of  e = eng_change - (T / 2) ln u  (flip <=> e >= 0; the f32
        // difference has the sign of the exact one, and exact equality gives +0), shifted into `keep` with one funnel shift per lane
        // (bit = 1: the lane does not flip) instead of a compare, a select and an OR.
        uint32_t keep = 0u;
#pragma unroll
        for (int g = 7; g >= 0; g--) {
            if (4 * g < nl) {
                uint32_t dr[4] = {0u, 0u, 0u, 0u};
                if (UPDATE) {
                    const u32x4 o = philox4x32_k(c, sweep_index, 8u * (uint32_t)w + (uint32_t)g, tagc, ks);
                    dr[0] = o.x >> 8; dr[1] = o.y >> 8; dr[2] = o.z >> 8; dr[3] = o.w >> 8;
                }
#pragma unroll
                for (int q = 3; q >= 0; q--) {
                    const int l = 4 * g + q, sh = 31 - l;
                    // sweep.rs:10-17: forward then backward per direction; a product with a +-1 spin is the coupling with its sign
                    // bit flipped.  The words were XORed with the site's own word, so the sum is s_i h (negation commutes with
                    // every rounding of the sum): -eng_change of sweep.rs:43-44 without a per-lane sign flip of its own.
                    float sh_ = __uint_as_float(__float_as_uint(Jf[0]) ^ ((F[0] << sh) & 0x80000000u));
                    sh_ = __fadd_rn(sh_, __uint_as_float(__float_as_uint(Jb[0]) ^ ((B[0] << sh) & 0x80000000u)));
#pragma unroll
                    for (int kk = 1; kk < Z; kk++) {
                        sh_ = __fadd_rn(sh_, __uint_as_float(__float_as_uint(Jf[kk]) ^ ((F[kk] << sh) & 0x80000000u)));
                        sh_ = __fadd_rn(sh_, __uint_as_float(__float_as_uint(Jb[kk]) ^ ((B[kk] << sh) & 0x80000000u)));
                    }
                    uint32_t ebits = 0x80000000u;  // no update: every lane keeps its spin
                    if (UPDATE) {
                        float thr;
                        if (EXACT) {  // host-libm tables: bit-identical to a host replay
                            thr = __fmul_rn(ht_sm[l], GIBBS ? m.glogtab[dr[q]] : m.logtab[dr[q]]);
                        } else {      // ln u = ln 2 * (log2(draw) - 24); Gibbs: ln(u / (1 - u)) = ln 2 * (log2(draw) - log2(2^24 - draw))
                            const float2 kt = kt_sm[l];
                            const float a = sw_lg2((float)dr[q]);
                            thr = GIBBS ? kt.x * (a - sw_lg2((float)(16777216u - dr[q]))) : fmaf(kt.x, a, kt.y);
                        }
                        ebits = __float_as_uint(__fadd_rn(-sh_, -thr));  // eng_change - (T / 2) ln u: sweep.rs:256 / 279-282
                        keep = __funnelshift_l(ebits, keep, 1);
                    }
                    // s h after the update: +s h where the lane keeps its spin, -s h where it flips
                    if (EACC) eacc[l] += __float_as_int(fmaf(sh_, __uint_as_float(__float_as_uint(escale) ^ (~ebits & 0x80000000u)), 12582912.0f));
                }
            } else if (UPDATE) {
                keep <<= 4;
            }
        }
        const uint32_t flips = ~keep;
        if (UPDATE) W[i] = C ^ (flips & vmask);
    }
    if (EACC) {
#pragma unroll
        for (int l = 0; l < 32; l++) {
            const int ws = __reduce_add_sync(0xFFFFFFFFu, (int)((uint32_t)eacc[l] - nterms * 0x4B400000u));  // minus the addends' bits
            if ((tid & 31) == l) atomicAdd(&blk_sm[l], ws);
        }
        __syncthreads();
        long long *acc = sv.acc_e + ((size_t)d * sv.SW + w) * 32;
        if (tid < nl) atomicAdd((unsigned long long *)&acc[tid], (unsigned long long)(long long)blk_sm[tid]);
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(&sv.arrive_e[blockIdx.x], 1u) == gridDim.y - 1 ? 1 : 0;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        if (tid < nl) {
            const long long e_tot = (long long)atomicExch((unsigned long long *)&acc[tid], 0ull);
            m.energies[d * m.S + 32 * w + tid] = __fdiv_rn((float)((double)e_tot / (double)escale), (float)m.N);
        }
        if (tid == 0) sv.arrive_e[blockIdx.x] = 0u;
    }
}

// which temperature every system sits at (parallel.rs:27-33, realization.rs:166): rebuilt from system_ids before the colour passes of
// a launch sequence, so that a colour-pass block reads its 32 lanes' thresholds with one load.  One thread per (realization, slot).
__global__ void __launch_bounds__(256) swords_lane_temps_kernel(ModelView m, SWordsView sv) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= m.D * m.S) return;
    const int64_t d = idx / m.S;
    const int slot = (int)(idx - d * m.S);
    const int sys = m.system_ids[idx];
    const float ht = __fdiv_rn(m.temps[slot % m.T], 2.0f), kt = ht * 0.693147180559945f;
    sv.lane_t[(size_t)d * sv.SW * 32 + sys] = make_float4(ht, kt, -24.0f * kt, 0.0f);
}

// words -> transposed view (+ magnetisation sums).  grid = (D * SW, ceil(N / 32 / SW_THREADS)); a thread owns 32 consecutive sites.
__global__ void __launch_bounds__(SW_THREADS) swords_transpose_kernel(ModelView m, SWordsView sv, int want_mags) {
    __shared__ int blk_sm[32];
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const int64_t d = blockIdx.x / sv.SW;
    const int w = (int)(blockIdx.x % sv.SW);
    const int nl = min(32, m.S - 32 * w);
    const uint32_t nq = (uint32_t)(m.N / 32), q = blockIdx.y * SW_THREADS + tid;
    if (tid < 32) blk_sm[tid] = 0;
    __syncthreads();
    uint32_t a[32];
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = 0u;
    if (q < nq) {
        const uint4 *src = reinterpret_cast<const uint4 *>(sv.words + ((size_t)d * sv.SW + w) * m.N + (size_t)q * 32);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint4 t = src[k];
            a[4 * k] = t.x; a[4 * k + 1] = t.y; a[4 * k + 2] = t.z; a[4 * k + 3] = t.w;
        }
        sw_transpose32(a);
        uint32_t *dst = sv.tbits + ((size_t)d * m.S + 32 * (size_t)w) * nq + q;
#pragma unroll
        for (int l = 0; l < 32; l++)
            if (l < nl) dst[(size_t)l * nq] = a[l];
    }
    if (!want_mags) return;
#pragma unroll
    for (int l = 0; l < 32; l++) {
        const int ws = __reduce_add_sync(0xFFFFFFFFu, __popc(a[l]));
        if ((tid & 31) == l) atomicAdd(&blk_sm[l], ws);
    }
    __syncthreads();
    long long *acc = sv.acc_m + ((size_t)d * sv.SW + w) * 32;
    if (tid < nl) atomicAdd((unsigned long long *)&acc[tid], (unsigned long long)(long long)blk_sm[tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(&sv.arrive_m[blockIdx.x], 1u) == gridDim.y - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid < nl) {
        const long long dn = (long long)atomicExch((unsigned long long *)&acc[tid], 0ull);
        m.mags[d * m.S + 32 * w + tid] = m.N - 2 * dn;  // energy.rs:85-90
    }
    if (tid == 0) sv.arrive_m[blockIdx.x] = 0u;
}

// replica-pair dots on the transposed view (overlap.rs:259-281).  One WARP per pair (d, p, t): the XOR of the two replicas' bit rows
// goes to the warp's own shared-memory region (N / 8 bytes), then q and q_link are popcounts of whole row words against the
// neighbour rows' words — no block-wide barrier.  grid = ceil(D * P * T / warps per CTA), dynamic smem = warps * N / 8 bytes.
template <int Z>
__global__ void __launch_bounds__(256) swords_overlap_kernel(ModelView m, RowsView v, SWordsView sv, long long *dot_spin, long long *dot_link) {
    extern __shared__ uint32_t sw_x_sm[];  // [warps][N / 32] bit = 1 where the two replicas differ
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int64_t idx = (int64_t)blockIdx.x * wpc + wid;  // (d * P + p) * T + t
    if (idx >= m.D * m.P * m.T) return;
    const int t = (int)(idx % m.T);
    const int p = (int)((idx / m.T) % m.P);
    const int64_t d = idx / ((int64_t)m.T * m.P);
    const int sa = m.system_ids[d * m.S + (2 * p) * m.T + t];
    const int sb = m.system_ids[d * m.S + (2 * p + 1) * m.T + t];
    const uint32_t nq = (uint32_t)(m.N / 32);
    const uint32_t *a = sv.tbits + ((size_t)d * m.S + sa) * nq, *b = sv.tbits + ((size_t)d * m.S + sb) * nq;
    uint32_t *X = sw_x_sm + (size_t)wid * nq;
    for (uint32_t q = lane; q < nq; q += 32) X[q] = a[q] ^ b[q];
    __syncwarp();
    int neg_q = 0, neg_l = 0;  // sites with q_i = -1, links with q_i q_j = -1 (each below 2^31: N < 2^28 at eligibility)
    if (v.L % 32 == 0) {
        const int Wr = v.L / 32;  // words per row
        for (uint32_t q = lane; q < nq; q += 32) {
            const uint32_t r = q / (uint32_t)Wr;
            const int wj = (int)(q - r * (uint32_t)Wr);
            const uint32_t Xq = X[q];
            neg_q += __popc(Xq);
#pragma unroll
            for (int kk = 0; kk < Z; kk++) {
                const uint32_t *row = X + (size_t)v.nbr_row[((size_t)r * Z + kk) * 2] * Wr;
                const int dl = v.dl[kk];
                uint32_t Xn = row[wj];
                if (dl > 0) Xn = (Xn >> 1) | (row[wj + 1 == Wr ? 0 : wj + 1] << 31);
                else if (dl < 0) Xn = (Xn << 1) | (row[wj ? wj - 1 : Wr - 1] >> 31);
                neg_l += __popc(Xq ^ Xn);
            }
        }
    } else {  // rows shorter than a word (L = 8, 16, 24): a row is L bits starting at bit r * L, possibly across two words
        const int L = v.L;
        const uint32_t rmask = (1u << L) - 1u;
        auto row_bits = [&](const uint32_t r) {
            const uint32_t start = r * (uint32_t)L, w0 = start >> 5, sh = start & 31u;
            const uint32_t lo = X[w0], hi = sh + (uint32_t)L > 32u ? X[w0 + 1] : 0u;
            return __funnelshift_r(lo, hi, sh) & rmask;
        };
        for (uint32_t r = lane; r < (uint32_t)v.n_rows; r += 32) {
            const uint32_t Xr = row_bits(r);
            neg_q += __popc(Xr);
#pragma unroll
            for (int kk = 0; kk < Z; kk++) {
                uint32_t Xn = row_bits(v.nbr_row[((size_t)r * Z + kk) * 2]);
                const int dl = v.dl[kk];
                if (dl > 0) Xn = ((Xn >> 1) | (Xn << (L - 1))) & rmask;       // site x looks at x + 1 (periodic in the row)
                else if (dl < 0) Xn = ((Xn << 1) | (Xn >> (L - 1))) & rmask;
                neg_l += __popc(Xr ^ Xn);
            }
        }
    }
    neg_q = __reduce_add_sync(0xFFFFFFFFu, neg_q);
    neg_l = __reduce_add_sync(0xFFFFFFFFu, neg_l);
    if (lane == 0) {
        dot_spin[idx] = m.N - 2ll * neg_q;
        dot_link[idx] = (long long)Z * m.N - 2ll * neg_l;
    }
}

// int8 view <-> words; dir 0: pack (int8 -> words), 1: unpack.  grid = (D * SW, ceil(N / 256))
__global__ void __launch_bounds__(256) swords_convert_kernel(ModelView m, SWordsView sv, int dir) {
    const int64_t d = blockIdx.x / sv.SW;
    const int w = (int)(blockIdx.x % sv.SW);
    const int nl = min(32, m.S - 32 * w);
    const int64_t i = (int64_t)blockIdx.y * 256 + threadIdx.x;
    if (i >= m.N) return;
    uint32_t *word = sv.words + ((size_t)d * sv.SW + w) * m.N + i;
    int8_t *sp = m.spins + ((size_t)d * m.S + 32 * (size_t)w) * m.N + i;
    if (dir == 0) {
        uint32_t x = 0u;
        for (int l = 0; l < nl; l++) x |= (uint32_t)(sp[(size_t)l * m.N] < 0) << l;
        *word = x;
    } else {
        const uint32_t x = *word;
        for (int l = 0; l < nl; l++) sp[(size_t)l * m.N] = (x >> l) & 1u ? (int8_t)-1 : (int8_t)1;
    }
}

#endif  // __CUDACC__

}  // namespace pp
