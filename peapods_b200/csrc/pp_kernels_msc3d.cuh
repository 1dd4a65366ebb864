// pp_kernels_msc3d.cuh — the headline kernel: multispin-coded Metropolis / heat-bath sweeps on 3-D
// hypercubic lattices with stride geometry (no neighbour tables), fused with the per-sweep energy,
// magnetisation and replica-overlap reductions.
//
// Replaces, for this layout, the reference loops
//   metropolis_sweep / gibbs_sweep          spin-sim/src/mcmc/sweep.rs:220-284 (+ :8-19 local field, :170-185 lookup rule)
//   compute_energies_and_magnetizations_into spin-sim/src/spins/energy.rs:59-110
//   OverlapAccum::collect (integer dots)     spin-sim/src/statistics/overlap.rs:259-281
//
// Storage (see pp_plan.h "compact" order): one u32 word = the same site of 32 disorder realizations
// (bit = 1: spin -1).  Inside a system the N words are checkerboard-compacted:
//   index(c, x0, x1, j) = c*N/2 + (x0*L1 + x1)*(L2/2) + j      colour c = (x0+x1+x2)&1, j = x2>>1
// so the four same-colour sites of an aligned 8-site segment of a row are ONE 128-bit word group and
// every neighbour of such a quad is again a 128-bit group of the other colour (plus one extra word for
// the +-x2 neighbour that crosses the segment).  A CTA owns temperature slot t of word group g for all
// R replicas: the R*N words live in shared memory for the whole launch (one bulk-async load, one
// bulk-async store per system = exactly the algorithmic 2 x 4 B per word); two such slots share one CTA
// (two independent 256-thread halves) together with one shared-memory copy of the group's coupling sign
// words, which are read once per quad and reused by the R replicas.
//
// Per word update (Metropolis): 6 three-input XORs (bond words), two bit-sliced full adders and five
// LOP3s give the masks [unsat >= 1], [>= 2], [>= 3]; the 24-bit draw (one Philox call per quad
// and replica, RNG-SPEC TAG_SWEEP_MSC) selects which of them is the flip mask:
//   flip lane l  <=>  draw < table[t][2*unsat_l]   (sweep.rs:182-184 with ec + 2z' = 2*unsat).
#pragma once
#ifndef PP_M3_VARIANT
#define PP_M3_VARIANT 2
#endif
#include <string>
#include <vector>

#include "../../include/peapods_b200.h"
#include "pp_device.cuh"
#include "pp_kernels_stats.cuh"
#include "pp_plan.h"

namespace pp {

// ------------------------------------------------------------------------------------------------
// host side: which lattices the kernel takes, and the per-item offset table
struct Msc3dPlan {
    bool ok = false;
    int L0 = 0, L1 = 0, L2 = 0, LXH = 0, QPR = 0, n_items = 0;
    std::vector<uint16_t> items;  // [n_items][8]: self, zp, zm, yp, ym, eL, eR, parity (word offsets inside one colour half)
};

inline Msc3dPlan msc3d_plan(const LatticePlan &p) {
    Msc3dPlan q;
    if (!(p.hypercubic && p.n_dims == 3 && p.compact && p.linear_colouring && p.colour_mod == 2)) return q;
    if (!(p.colour_coef[0] == 1 && p.colour_coef[1] == 1 && p.colour_coef[2] == 1)) return q;
    q.L0 = (int)p.shape[0]; q.L1 = (int)p.shape[1]; q.L2 = (int)p.shape[2];
    if (q.L0 % 2 || q.L1 % 2 || q.L2 % 8) return q;
    if (p.n_spins / 2 > 65535) return q;
    q.LXH = q.L2 / 2;
    q.QPR = q.LXH / 4;
    q.n_items = q.L0 * q.L1 * q.QPR;
    q.items.resize((size_t)q.n_items * 8);
    // Item order: all segments of rows with (x0 + x1) even first, then the odd ones, so that the 32 items a warp takes
    // together share the x2-parity of their sites (no divergence on it); inside a parity class walk (x0>>1, x1, h), which
    // alternates between the two x0 planes of a pair and keeps the 128-bit shared-memory accesses of a quarter-warp
    // on distinct banks for the usual power-of-two extents.
    int it = 0;
    for (int par = 0; par < 2; par++)
        for (int zb = 0; zb < q.L0 / 2; zb++)
            for (int x1 = 0; x1 < q.L1; x1++)
                for (int h = 0; h < q.QPR; h++, it++) {
                    const int x0 = 2 * zb + ((x1 + par) & 1);
                    auto at = [&](int a, int b, int j) { return (uint16_t)((a * q.L1 + b) * q.LXH + j); };
                    uint16_t *d = &q.items[(size_t)it * 8];
                    d[0] = at(x0, x1, 4 * h);
                    d[1] = at((x0 + 1) % q.L0, x1, 4 * h);
                    d[2] = at((x0 + q.L0 - 1) % q.L0, x1, 4 * h);
                    d[3] = at(x0, (x1 + 1) % q.L1, 4 * h);
                    d[4] = at(x0, (x1 + q.L1 - 1) % q.L1, 4 * h);
                    d[5] = at(x0, x1, (4 * h + q.LXH - 1) % q.LXH);
                    d[6] = at(x0, x1, (4 * h + 4) % q.LXH);
                    d[7] = (uint16_t)par;
                }
    q.ok = true;
    return q;
}

struct Msc3dView {
    const uint4 *items;  // [n_items] packed 8 x u16
    int n_items;
    uint32_t N, N2;
    uint32_t one[3];     // the constant 1, three times, opaque to the compiler (multipliers of the IMAD.WIDE compares, see lt_mask;
                         // distinct operands keep ptxas from sharing one product and doing the 64-bit adds on the ALU pipe)
};

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// 1-D bulk-async copy global -> shared (TMA engine), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void half_barrier(int half, int nthreads) {  // named barrier of one half of the CTA
    asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "r"(nthreads) : "memory");
}

// bit-sliced vertical counter: plane b holds bit b of 32 independent per-lane counts
template <int K>
struct VAcc {
    uint32_t p[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int b = 0; b < K; b++) p[b] = 0u;
    }
    // add one word of weight 2^LVL, ripple to the top
    template <int LVL>
    __device__ __forceinline__ void add1(uint32_t x) {
#pragma unroll
        for (int b = LVL; b < K; b++) {
            const uint32_t t = p[b] & x;
            p[b] ^= x;
            x = t;
        }
    }
    // carry-save: add two words of weight 2^LVL into plane LVL, return the carry word (weight 2^(LVL+1))
    template <int LVL>
    __device__ __forceinline__ uint32_t csa(uint32_t a, uint32_t b) {
        const uint32_t c = maj3(p[LVL], a, b);
        p[LVL] = xor3(p[LVL], a, b);
        return c;
    }
    template <int LVL>
    __device__ __forceinline__ uint32_t half(uint32_t a) {
        const uint32_t c = p[LVL] & a;
        p[LVL] ^= a;
        return c;
    }
    // add a KI-plane bit-sliced number (ripple carry)
    template <int KI>
    __device__ __forceinline__ void add_planes(const uint32_t *x) {
        uint32_t c = 0u;
#pragma unroll
        for (int b = 0; b < K; b++) {
            if (b < KI) {
                const uint32_t sum = xor3(p[b], x[b], c);
                c = maj3(p[b], x[b], c);
                p[b] = sum;
            } else {
                const uint32_t t = p[b] & c;
                p[b] ^= c;
                c = t;
            }
        }
    }
    // eight words of weight 1
    __device__ __forceinline__ void add8(const uint32_t *w) {
        const uint32_t k0 = csa<0>(w[0], w[1]), k1 = csa<0>(w[2], w[3]), k2 = csa<0>(w[4], w[5]), k3 = csa<0>(w[6], w[7]);
        const uint32_t m0 = csa<1>(k0, k1), m1 = csa<1>(k2, k3);
        const uint32_t n0 = csa<2>(m0, m1);
        add1<3>(n0);
    }
    // eight words of weight 1 (s) and eight of weight 2 (c)
    __device__ __forceinline__ void add8_8(const uint32_t *s, const uint32_t *c) {
        const uint32_t k0 = csa<0>(s[0], s[1]), k1 = csa<0>(s[2], s[3]), k2 = csa<0>(s[4], s[5]), k3 = csa<0>(s[6], s[7]);
        const uint32_t m0 = csa<1>(c[0], c[1]), m1 = csa<1>(c[2], c[3]), m2 = csa<1>(c[4], c[5]), m3 = csa<1>(c[6], c[7]);
        const uint32_t m4 = csa<1>(k0, k1), m5 = csa<1>(k2, k3);
        const uint32_t n0 = csa<2>(m0, m1), n1 = csa<2>(m2, m3), n2 = csa<2>(m4, m5);
        const uint32_t r0 = csa<3>(n0, n1), r1 = half<3>(n2);
        const uint32_t q0 = csa<4>(r0, r1);
        add1<5>(q0);
    }
};

// Sum the K-plane counters of the 32 threads of a warp (bit-sliced butterfly); thread l returns the total of lane l.
template <int K>
__device__ __forceinline__ uint32_t warp_lane_total(const VAcc<K> &v) {
    constexpr int KW = K + 5;
    uint32_t a[KW];
#pragma unroll
    for (int b = 0; b < KW; b++) a[b] = b < K ? v.p[b] : 0u;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const int o = 16 >> r;
        uint32_t carry = 0;
#pragma unroll
        for (int b = 0; b < KW; b++) {
            if (b > K + r) break;  // planes above K+r are still zero in round r
            const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, a[b], o);
            const uint32_t s = xor3(a[b], y, carry);
            carry = maj3(a[b], y, carry);
            a[b] = s;
        }
    }
    const int lane = threadIdx.x & 31;
    uint32_t total = 0;
#pragma unroll
    for (int b = 0; b < KW; b++) total += ((a[b] >> lane) & 1u) << b;
    return total;
}

__device__ __forceinline__ void unpack_item(const uint4 d, uint32_t &self, uint32_t &zp, uint32_t &zm, uint32_t &yp,
                                            uint32_t &ym, uint32_t &eL, uint32_t &eR, uint32_t &par) {
    self = d.x & 0xFFFFu; zp = d.x >> 16;
    zm = d.y & 0xFFFFu;   yp = d.y >> 16;
    ym = d.z & 0xFFFFu;   eL = d.z >> 16;
    eR = d.w & 0xFFFFu;   par = d.w >> 16;
}

// Word-wide comparison on the FMA pipe: all-ones if raw < thr else 0, as the high half of raw * 1 + (2^64 - thr)
// (one IMAD.WIDE.U32; the ALU pipe, which the bit-sliced logic saturates, is not touched).  `one` must be 1.
__device__ __forceinline__ uint32_t lt_mask(const uint32_t raw, const uint32_t one, const uint64_t neg_thr) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(raw), "r"(one), "l"(neg_thr));
    return (uint32_t)(r >> 32);
}

__device__ __forceinline__ uint4 lds4(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void to_arr(const uint4 v, uint32_t *a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }

constexpr int MSC3D_KE = 10;  // planes of the per-thread unsatisfied-bond counters (3*sites_per_thread < 1016)
constexpr int MSC3D_KM = 9;   // planes of the per-thread down-spin / q counters     (sites_per_thread   < 504)
constexpr int MSC3D_NTH = 256;  // threads per half
constexpr int MSC3D_NBAR = 18;  // mbarriers in shared memory: [0] couplings (+ item table), [1 + half * 8 + r] the spin words of replica r of a half

// One quad (four same-colour sites of one row segment) of all RPC replicas.  P = the quad's sites sit at x2 = 2j + P.
//   sp: the half's [RPC][N] words; J: [3][N] coupling sign words in shared memory (all-zero words for a ferromagnet)
constexpr int MSC3D_KS = 6;  // planes of the in-sweep unsatisfied-bond counters (two quads per thread: 2 * 4 * 6 = 48)

// Sum of eight words of weight 1 (w1) and eight of weight 2 (w2) as a 5-plane number (<= 24): carry-save tree, 28 LOP3.
__device__ __forceinline__ void tree_8_8(const uint32_t *w1, const uint32_t *w2, uint32_t *t) {
    const uint32_t S0 = xor3(w1[0], w1[1], w1[2]), C0 = maj3(w1[0], w1[1], w1[2]);
    const uint32_t S1 = xor3(w1[3], w1[4], w1[5]), C1 = maj3(w1[3], w1[4], w1[5]);
    const uint32_t S2 = xor3(S0, S1, w1[6]), C2 = maj3(S0, S1, w1[6]);
    t[0] = S2 ^ w1[7];
    const uint32_t C3 = S2 & w1[7];
    const uint32_t S3 = xor3(w2[0], w2[1], w2[2]), D0 = maj3(w2[0], w2[1], w2[2]);
    const uint32_t S4 = xor3(w2[3], w2[4], w2[5]), D1 = maj3(w2[3], w2[4], w2[5]);
    const uint32_t S5 = xor3(w2[6], w2[7], C0), D2 = maj3(w2[6], w2[7], C0);
    const uint32_t S6 = xor3(C1, C2, C3), D3 = maj3(C1, C2, C3);
    const uint32_t S7 = xor3(S3, S4, S5), D4 = maj3(S3, S4, S5);
    t[1] = S7 ^ S6;
    const uint32_t D5 = S7 & S6;
    const uint32_t S8 = xor3(D0, D1, D2), E0 = maj3(D0, D1, D2);
    const uint32_t S9 = xor3(D3, D4, D5), E1 = maj3(D3, D4, D5);
    t[2] = S8 ^ S9;
    const uint32_t E2 = S8 & S9;
    t[3] = xor3(E0, E1, E2);
    t[4] = maj3(E0, E1, E2);
}

// ACC: also count, per lane, the unsatisfied bonds of the updated sites AFTER the update into ea[r] (a flip inverts the six
// bond words of its site, so the two adder outputs of each half are simply XORed with the flip mask).  On a bipartite
// lattice the six bonds of the colour-1 sites are all bonds, each once: the colour-1 pass of the last sweep yields the
// energy of energy.rs:99-108 without another pass over the spins.
//
// P (runtime, warp-uniform: the item order keeps the quads of one x2-parity together): the quad's sites sit at x2 = 2j + P.
// Their two x2 neighbours are the aligned quad O of the other colour (the left neighbours if P, the right ones if not) and
// that quad shifted by one word towards the other side, completed by the row's wrap word.  Both bonds enter the same
// adder, so only the pairing (spin words, coupling words) matters; parity is a matter of word offsets, not of code.
template <int RPC, bool METRO, bool ACC>
__device__ __forceinline__ void msc3d_sweep_item(uint32_t *sp, const uint32_t *J, const uint32_t N, const uint32_t so,
                                                 const uint32_t oo, const uint4 desc, const bool P, const uint32_t (&thr)[7],
                                                 const uint64_t (&nthr)[3], const uint32_t (&one)[3], const uint32_t sweep, const uint32_t pos0, const uint32_t pos_stride,
                                                 const uint32_t tag, const uint32_t k0, const uint32_t k1,
                                                 VAcc<MSC3D_KS> *ea) {
    uint32_t self, zp, zm, yp, ym, eL, eR, par;
    unpack_item(desc, self, zp, zm, yp, ym, eL, eR, par);
    (void)par;
    // word offsets of the shifted quad inside the other colour half
    const uint32_t sh0 = P ? self + 1 : eL, sh1 = self + (P ? 2u : 0u), sh2 = self + (P ? 3u : 1u), sh3 = P ? eR : self + 2;
    // coupling sign words (bit = 1: J = -1); bond (i, d) is stored at its lower site i: the forward bonds at the quad itself,
    // the backward bonds at the backward neighbours
    uint32_t Jf0[4], Jf1[4], Jb0[4], Jb1[4], JO[4], JS[4];
    to_arr(lds4(J + 0 * N + so + self), Jf0);
    to_arr(lds4(J + 1 * N + so + self), Jf1);
    to_arr(lds4(J + 0 * N + oo + zm), Jb0);
    to_arr(lds4(J + 1 * N + oo + ym), Jb1);
    {
        const uint32_t *J2 = J + 2 * N;
        to_arr(lds4(J2 + (P ? oo : so) + self), JO);  // P: O = left neighbours, bond stored there; else O = right, stored here
        const uint32_t *JSb = J2 + (P ? so : oo);     // P: shifted = right neighbours, bonds stored at the quad itself
        JS[0] = JSb[P ? self : sh0]; JS[1] = JSb[P ? self + 1 : sh1]; JS[2] = JSb[P ? self + 2 : sh2]; JS[3] = JSb[P ? self + 3 : sh3];
    }
#pragma unroll
    for (int r = 0; r < RPC; r++) {
        uint32_t *sys = sp + (size_t)r * N;
        const u32x4 rnd = philox4x32(self >> 2, sweep, pos0 + (uint32_t)r * pos_stride, tag, k0, k1);
        uint32_t s[4], zpv[4], zmv[4], ypv[4], ymv[4], xo[4], xs[4];
        to_arr(lds4(sys + so + self), s);
        to_arr(lds4(sys + oo + zp), zpv);
        to_arr(lds4(sys + oo + zm), zmv);
        to_arr(lds4(sys + oo + yp), ypv);
        to_arr(lds4(sys + oo + ym), ymv);
        to_arr(lds4(sys + oo + self), xo);
        xs[0] = sys[oo + sh0]; xs[1] = sys[oo + sh1]; xs[2] = sys[oo + sh2]; xs[3] = sys[oo + sh3];
        const uint32_t raw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
        uint32_t w1[8], w2[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b0 = xor3(s[j], zpv[j], Jf0[j]), b1 = xor3(s[j], zmv[j], Jb0[j]);
            const uint32_t b2 = xor3(s[j], ypv[j], Jf1[j]), b3 = xor3(s[j], ymv[j], Jb1[j]);
            const uint32_t b4 = xor3(s[j], xo[j], JO[j]), b5 = xor3(s[j], xs[j], JS[j]);
            const uint32_t s1 = xor3(b0, b1, b2), c1 = maj3(b0, b1, b2);
            const uint32_t s2 = xor3(b3, b4, b5), c2 = maj3(b3, b4, b5);
            const uint32_t kk = s1 & s2, oo2 = s1 | s2;
            uint32_t flip;
            if (METRO) {
                // thr[u] = count[u] << 8 with count[u] < 2^24 for unsat u = 0..2 and count = 2^24 (always accept,
                // sweep.rs:141-145) for u >= 3:  (raw >> 8) < count  <=>  raw < thr.  The events nest (p0 => p1 => p2), so
                // with x = s1 + s2, y = c1 + c2, L = #events:  flip <=> x + 2y + L >= 3.
#if PP_M3_VARIANT == 0
                const uint32_t ge1 = oo2 | c1 | c2, ge2 = kk | c1 | c2, ge3 = maj3(c1, c2, oo2);
                flip = ge3;
                if (raw[j] < thr[2]) flip = ge2;
                if (raw[j] < thr[1]) flip = ge1;
                if (raw[j] < thr[0]) flip = 0xFFFFFFFFu;
#elif PP_M3_VARIANT == 1
                const uint32_t M2 = lt_mask(raw[j], one[2], nthr[2]), M1 = lt_mask(raw[j], one[1], nthr[1]);
                const uint32_t M0 = lt_mask(raw[j], one[0], nthr[0]);
                const uint32_t a3 = maj3(c1, c2, oo2 | M2);  // y = 2, or y = 1 and x + L >= 1
                const uint32_t b3 = M0 | (oo2 & M1);         // L = 3, or x >= 1 and L >= 2
                flip = a3 | (kk & M2) | b3;                  // or x = 2 and L >= 1
#else
                const bool p2 = raw[j] < thr[2], p1 = raw[j] < thr[1], p0 = raw[j] < thr[0];
                flip = maj3(c1, c2, p2 ? 0xFFFFFFFFu : oo2);  // y = 2, or y = 1 and x + L >= 1
                if (p2) flip |= kk;                           // x = 2 and L >= 1
                if (p1) flip |= oo2;                          // x >= 1 and L >= 2
                if (p0) flip = 0xFFFFFFFFu;                   // L = 3
#endif
            } else {  // thr[u] = count[u]
                const uint32_t x0 = s1 ^ s2, y1 = xor3(c1, c2, kk), y2 = maj3(c1, c2, kk);
                const uint32_t draw = raw[j] >> 8;
                flip = 0u;
#pragma unroll
                for (int u = 0; u < 7; u++) {
                    const uint32_t eq = ((u & 1) ? x0 : ~x0) & ((u & 2) ? y1 : ~y1) & ((u & 4) ? y2 : ~y2);
                    if (draw < thr[u]) flip |= eq;
                }
            }
            s[j] ^= flip;
            if (ACC) {
                w1[2 * j] = s1 ^ flip; w1[2 * j + 1] = s2 ^ flip;
                w2[2 * j] = c1 ^ flip; w2[2 * j + 1] = c2 ^ flip;
            }
        }
        *reinterpret_cast<uint4 *>(sys + so + self) = make_uint4(s[0], s[1], s[2], s[3]);
        if (ACC) {
            uint32_t t5[5];
            tree_8_8(w1, w2, t5);
            ea[r].template add_planes<5>(t5);
        }
    }
}

// The 8 sites of one row segment of system A with their +x0, +x1, +x2 neighbours, as two groups of four:
//   X = the colour whose sites sit at odd x2 (colour 0 if PAR, else colour 1): its +x2 neighbours are the other colour's quad
//       shifted by one word (plus the row's wrap word eR);
//   Y = the other colour: its +x2 neighbours are X's own words.
// xo / yo = word offsets of the two colour halves (0 or N2).  Words 0..3 of each array = group X, 4..7 = group Y.
__device__ __forceinline__ void msc3d_load_segment(const uint32_t *A, const uint32_t xo, const uint32_t yo, const uint32_t self,
                                                   const uint32_t zp, const uint32_t yp, const uint32_t eR, uint32_t *a,
                                                   uint32_t *az, uint32_t *ay, uint32_t *ax) {
    to_arr(lds4(A + xo + self), a); to_arr(lds4(A + yo + self), a + 4);
    to_arr(lds4(A + yo + zp), az); to_arr(lds4(A + xo + zp), az + 4);
    to_arr(lds4(A + yo + yp), ay); to_arr(lds4(A + xo + yp), ay + 4);
    ax[0] = a[5]; ax[1] = a[6]; ax[2] = a[7]; ax[3] = A[yo + eR];
    ax[4] = a[0]; ax[5] = a[1]; ax[6] = a[2]; ax[7] = a[3];
}

__device__ __forceinline__ void msc3d_em_item(const uint32_t *A, const uint32_t *J, const uint32_t N, const uint32_t N2,
                                              const uint4 desc, const int want_energy, const int want_mags,
                                              VAcc<MSC3D_KE> &ve, VAcc<MSC3D_KM> &vm) {
    uint32_t self, zp, zm, yp, ym, eL, eR, par;
    unpack_item(desc, self, zp, zm, yp, ym, eL, eR, par);
    (void)zm; (void)ym; (void)eL;
    const uint32_t xo = (par & 1u) ? 0u : N2, yo = N2 - xo;
    uint32_t a[8], az[8], ay[8], ax[8];
    msc3d_load_segment(A, xo, yo, self, zp, yp, eR, a, az, ay, ax);
    if (want_mags) vm.add8(a);
    if (want_energy) {
        uint32_t j0[8], j1[8], j2[8];
        to_arr(lds4(J + 0 * N + xo + self), j0); to_arr(lds4(J + 0 * N + yo + self), j0 + 4);
        to_arr(lds4(J + 1 * N + xo + self), j1); to_arr(lds4(J + 1 * N + yo + self), j1 + 4);
        to_arr(lds4(J + 2 * N + xo + self), j2); to_arr(lds4(J + 2 * N + yo + self), j2 + 4);
        uint32_t sb[8], cb[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {  // energy.rs:99-107: forward bonds only, each bond once
            const uint32_t b0 = xor3(a[j], az[j], j0[j]), b1 = xor3(a[j], ay[j], j1[j]), b2 = xor3(a[j], ax[j], j2[j]);
            sb[j] = xor3(b0, b1, b2);
            cb[j] = maj3(b0, b1, b2);
        }
        ve.add8_8(sb, cb);
    }
}

__device__ __forceinline__ void msc3d_pair_item(const uint32_t *A, const uint32_t *B, const uint32_t N2, const uint4 desc,
                                                VAcc<MSC3D_KE> &vl, VAcc<MSC3D_KM> &vq) {
    uint32_t self, zp, zm, yp, ym, eL, eR, par;
    unpack_item(desc, self, zp, zm, yp, ym, eL, eR, par);
    (void)zm; (void)ym; (void)eL;
    const uint32_t xo = (par & 1u) ? 0u : N2, yo = N2 - xo;
    uint32_t a[8], az[8], ay[8], ax[8], b[8], bz[8], by[8], bx[8];
    msc3d_load_segment(A, xo, yo, self, zp, yp, eR, a, az, ay, ax);
    msc3d_load_segment(B, xo, yo, self, zp, yp, eR, b, bz, by, bx);
    uint32_t x[8], sb[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {  // overlap.rs:266-276: x = 1 where q_i = -1; link word = x_i ^ x_fwd
        x[j] = a[j] ^ b[j];
        const uint32_t l0 = xor3(x[j], az[j], bz[j]), l1 = xor3(x[j], ay[j], by[j]), l2 = xor3(x[j], ax[j], bx[j]);
        sb[j] = xor3(l0, l1, l2);
        cb[j] = maj3(l0, l1, l2);
    }
    vq.add8(x);
    vl.add8_8(sb, cb);
}

// the pair item of the in-sweep-energy path: also the down-spin counts of both replicas (their words are loaded anyway)
constexpr int MSC3D_KQ = 6;  // planes of the per-thread q / down-spin counters of that path (4 items * 8 sites = 32)
constexpr int MSC3D_KL = 7;  // planes of its link counters (3 * 32 = 96)
__device__ __forceinline__ void msc3d_pairm_item(const uint32_t *A, const uint32_t *B, const uint32_t N2, const uint4 desc,
                                                 VAcc<MSC3D_KL> &vl, VAcc<MSC3D_KQ> &vq, VAcc<MSC3D_KQ> &vma,
                                                 VAcc<MSC3D_KQ> &vmb) {
    uint32_t self, zp, zm, yp, ym, eL, eR, par;
    unpack_item(desc, self, zp, zm, yp, ym, eL, eR, par);
    (void)zm; (void)ym; (void)eL;
    const uint32_t xo = (par & 1u) ? 0u : N2, yo = N2 - xo;
    uint32_t a[8], az[8], ay[8], ax[8], b[8], bz[8], by[8], bx[8];
    msc3d_load_segment(A, xo, yo, self, zp, yp, eR, a, az, ay, ax);
    msc3d_load_segment(B, xo, yo, self, zp, yp, eR, b, bz, by, bx);
    vma.add8(a);
    vmb.add8(b);
    uint32_t x[8], sb[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        x[j] = a[j] ^ b[j];
        const uint32_t l0 = xor3(x[j], az[j], bz[j]), l1 = xor3(x[j], ay[j], by[j]), l2 = xor3(x[j], ax[j], bx[j]);
        sb[j] = xor3(l0, l1, l2);
        cb[j] = maj3(l0, l1, l2);
    }
    vq.add8(x);
    vl.add8_8(sb, cb);
}

// Sum the KI-plane counters that `nw` warps parked in shared memory ([plane][nw * 32] words at `src`) lane by lane and
// return this thread's lane total (thread l: realization l of the word group).
// One out-of-line copy per (planes, warps) combination; the ripple of the k-th addend stops where its carry can still reach.
template <int KI, int NW>
__device__ __noinline__ uint32_t merge_lane_total(const uint32_t *src, const int lane) {
    constexpr int LG = NW <= 1 ? 0 : NW <= 2 ? 1 : NW <= 4 ? 2 : 3;
    VAcc<KI + LG> acc;
#pragma unroll
    for (int b = 0; b < KI + LG; b++) acc.p[b] = b < KI ? src[(b * NW) * 32 + lane] : 0u;
#pragma unroll
    for (int k = 1; k < NW; k++) {
        const int top = KI + (k < 2 ? 1 : k < 4 ? 2 : 3);  // planes of the running sum after this addend
        uint32_t c = 0u;
#pragma unroll
        for (int b = 0; b < KI + LG; b++) {
            if (b < KI) {
                const uint32_t x = src[(b * NW + k) * 32 + lane];
                const uint32_t sum = xor3(acc.p[b], x, c);
                c = maj3(acc.p[b], x, c);
                acc.p[b] = sum;
            } else if (b < top) {
                const uint32_t t = acc.p[b] & c;
                acc.p[b] ^= c;
                c = t;
            }
        }
    }
    return warp_lane_total(acc);
}

#ifdef PP_M3_TIMING  // phase clocks of every CTA of the last launch (tools/m3_phases.py; never defined in the product build)
__device__ unsigned long long pp_m3_clk[8192 * 32];
#define M3_CLK(k) do { if (threadIdx.x == 0 && blockIdx.x < 8192) pp_m3_clk[blockIdx.x * 32 + (k)] = clock64(); } while (0)
#define M3_T(k) M3_CLK(16 + (k))
#else
#define M3_CLK(k) do { } while (0)
#define M3_T(k) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// grid.x = G * ceil(T / NH); block = NH * 256 threads.  Half h of CTA (g, tp) owns temperature slot t = tp*NH + h of
// word group g for all RPC replicas; the halves share the coupling words and the item table in shared memory and
// otherwise run independently (named barriers), so one half can stage data while the other computes.
// dynamic smem (words): [3N coupling words | 4*n_items item table | NH*RPC*N spins | 8 (mbarriers) | NH*512 scratch]
// NFIX > 0: the site count is this compile-time constant (replica / direction strides become immediates); 0: gv.N
template <int RPC, bool METRO, int NH, int NFIX = 0>
__global__ void __launch_bounds__(MSC3D_NTH *NH, NH == 1 ? 2 : 1)
msc3d_kernel(ModelView m, Msc3dView gv, StatsView st, uint32_t sweep_index, int n_sweeps, int want_energy, int want_mags,
             int want_overlap, int want_fold, int64_t group_offset, long long *dot_spin, long long *dot_link,
             uint32_t *words_out, const uint32_t *swap_mask, int pt_schedule, int pt_parity, int esw) {
    extern __shared__ __align__(128) uint32_t smem[];
    const uint32_t N = NFIX > 0 ? (uint32_t)NFIX : gv.N, N2 = NFIX > 0 ? (uint32_t)NFIX / 2 : gv.N2;
    uint32_t *Jsm = smem;
    // NH == 2: the item table is staged in shared memory and the reduction scratch has its own space.
    // NH == 1: two CTAs share an SM (the spins and coupling words of one CTA are half of its shared memory to the
    // byte), so the item table is read through L1/L2 with a one-item register prefetch and the reduction scratch
    // reuses the spin buffer once the epilogue has consumed it.
    uint4 *items_sm = reinterpret_cast<uint4 *>(smem + 3 * N);
    uint32_t *sp_all = NH == 2 ? reinterpret_cast<uint32_t *>(items_sm + gv.n_items) : smem + 3 * N;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sp_all + (size_t)NH * RPC * N);
    uint16_t *ids_all = reinterpret_cast<uint16_t *>(bars + MSC3D_NBAR);  // [NH][RPC][32] system ids of the CTA's realizations, fetched at the top
    uint32_t *red_all = reinterpret_cast<uint32_t *>(ids_all + NH * 4 * 32);
    auto item_at = [&](int it) { return NH == 2 ? items_sm[it] : __ldg(gv.items + it); };
    const int tid = threadIdx.x;
    const int half = NH == 1 ? 0 : tid / MSC3D_NTH, ht = tid - half * MSC3D_NTH;
    const int TP = (m.T + NH - 1) / NH;
    const int64_t g = blockIdx.x / TP;
    const int t = (int)(blockIdx.x % TP) * NH + half;
    const uint32_t bytes = N * 4u;
    uint32_t *sp = sp_all + (size_t)half * RPC * N;
    uint32_t *red = NH == 2 ? red_all + half * 512 : sp;
    const uint32_t *J = Jsm;
#ifdef PP_M3_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 8192) {
        unsigned long long gt; uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        pp_m3_clk[blockIdx.x * 32 + 10] = gt; pp_m3_clk[blockIdx.x * 32 + 11] = smid;
    }
#endif
    M3_CLK(0);

    // ---- the exchange masks and thresholds first: everything after the bulk copies hangs on them (masks -> gather loads), and a
    // load that queues behind 112 KB of bulk requests comes back a microsecond later
    const bool live = t < m.T;
    // Spin words: m.words is the input buffer, words_out the output buffer (ping-pong, so that a CTA may still read its
    // neighbours' pre-sweep words while they are being rewritten).  When a parallel-tempering event is pending
    // (swap_mask != nullptr) the lanes that crossed an edge are gathered from the neighbouring slots while loading:
    //   single_random_edge: out = a ^ ((a ^ w[t-1]) & mask[t-1]) ^ ((a ^ w[t+1]) & mask[t])   (a lane attempts one edge)
    //   full_ladder:        the same exchange applied for the edges of the first parity, then of the second
    // (mcmc/tempering.rs:20-70: labels move there, lanes move here).  Systems without a crossing lane take the bulk path.
    uint32_t bulk_mask = 0;  // bit r: replica r is staged by a bulk-async copy
    uint32_t mk[RPC][3];     // schedule 0: {mask[t-1], mask[t], -}; schedule 1: {m1(t), m2(t), m1(t2)}
    int t1 = -1, t2 = -1, t21 = -1;  // full ladder: first-pass partner of t, second-pass partner of t, first-pass partner of t2
    // The mask words are loaded unconditionally from clamped edge indices, all loads back to back into their own registers, and
    // only looked at after the bulk copies have been issued (a branch per replica and side made every load wait for the one before).
    int edge[3] = {0, 0, 0};
    bool edge_on[3] = {false, false, false};
    const bool pend = swap_mask != nullptr && live && m.T > 1;
    if (pend) {
        auto partner = [&](int x, int q) {  // slot exchanging with x over an edge of parity q (edges q, q+2, ...), or -1
            if (x >= q && ((x - q) & 1) == 0 && x + 1 < m.T) return x + 1;
            if (x - 1 >= q && ((x - 1 - q) & 1) == 0) return x - 1;
            return -1;
        };
        if (pt_schedule == 0) {
            edge_on[0] = t > 0;       edge[0] = max(t - 1, 0);
            edge_on[1] = t + 1 < m.T; edge[1] = min(t, m.T - 2);
        } else {
            t1 = partner(t, pt_parity);
            t2 = partner(t, 1 - pt_parity);
            t21 = t2 >= 0 ? partner(t2, pt_parity) : -1;
            edge_on[0] = t1 >= 0;  edge[0] = t1 >= 0 ? min(t, t1) : 0;
            edge_on[1] = t2 >= 0;  edge[1] = t2 >= 0 ? min(t, t2) : 0;
            edge_on[2] = t21 >= 0; edge[2] = t21 >= 0 ? min(t2, t21) : 0;
        }
    }
    // One load per lane fetches every small word the CTA needs (lanes 0..23: mask k of replica r at lane r + 8k; lanes 24..30: the
    // acceptance thresholds count[u] = table[t][2u], sweep.rs:162-166, index ec + 2z' = 2*unsat); the words are handed round
    // by shuffles after the block-wide barrier below.  (Uniform loads per replica made the compiler move each word to a uniform
    // register straight away: one memory round trip per replica instead of one.)
    uint32_t small_word = 0u;
    {
        const int sl = ht & 31, sr = sl & 7, sk = sl >> 3;
        const uint32_t *addr = nullptr;
        if (sk < 3) {
            if (pend && sr < RPC)
                addr = swap_mask + (g * m.R + sr) * (int64_t)(m.T - 1) + (sk == 0 ? edge[0] : sk == 1 ? edge[1] : edge[2]);
        } else if (sl < 31 && live) {
            addr = m.lut + t * 13 + 2 * (sl - 24);
        }
        if (addr) small_word = __ldg(addr);
    }
    // ---- stage in: bulk-async copies (TMA engine), completion on mbarriers.  Issued right behind the one small load, each by the thread
    // that initialises its barrier (no block-wide wait in front of the first byte): thread 0 the couplings (+ item table),
    // lane 0 of warp (r + 1) & 7 of a half the spin words of replica r.
    static_assert(RPC <= 8, "one spin barrier per replica and half");
    auto spin_bar = [&](int r) { return &bars[1 + half * 8 + r]; };
    auto wait_spins = [&]() {
#pragma unroll
        for (int r = 0; r < RPC; r++) mbar_wait(spin_bar(r), 0);
    };
    if ((tid & 31) == 0) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&bars[0], 3u * bytes + (NH == 2 ? (uint32_t)gv.n_items * 16u : 0u));
            bulk_g2s(Jsm, m.Jw + g * 3 * (int64_t)N, 3u * bytes, &bars[0]);
            if (NH == 2) bulk_g2s(items_sm, gv.items, (uint32_t)gv.n_items * 16u, &bars[0]);
        }
        if (t < m.T) {
#pragma unroll
            for (int r = 0; r < RPC; r++)
                if (((r + 1) & 7) == (ht >> 5)) {
                    mbar_init(spin_bar(r), 1);
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    mbar_expect_tx(spin_bar(r), bytes);
                    bulk_g2s(sp + (size_t)r * N, m.words + ((g * m.S + (int64_t)r * m.T + t) * (int64_t)N), bytes, spin_bar(r));
                }
        }
    }
    M3_T(2);
    // The epilogue's last step addresses per-system outputs through system_ids: fetched now (the labels do not move during a
    // launch), parked in shared memory once the stage-in has waited anyway, so that the tail does not walk RPC cold loads.
    uint16_t *ids = ids_all + half * 4 * 32;
    int my_ids[RPC <= 4 ? RPC : 1];
    const bool ids_ok = RPC <= 4 && m.S <= 65535;
    const bool ids_pre = ids_ok && live && ht < 32 && g * 32 + ht < m.D;
    if (ids_pre) {
#pragma unroll
        for (int r = 0; r < (RPC <= 4 ? RPC : 1); r++) my_ids[r] = __ldg(m.system_ids + (g * 32 + ht) * m.S + r * m.T + t);
    }
    M3_T(0);
    __syncthreads();  // the barriers initialised above are visible to every waiter
    M3_CLK(15);
    if (t >= m.T) return;  // odd T: the last CTA of a group has an idle half
    M3_T(3);
    uint32_t thr[7];
#pragma unroll
    for (int u = 0; u < 7; u++) {
        const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, small_word, 24 + u);
        thr[u] = METRO ? (cnt << 8) : cnt;
    }
#pragma unroll
    for (int r = 0; r < RPC; r++) {
#pragma unroll
        for (int k = 0; k < 3; k++) mk[r][k] = __shfl_sync(0xFFFFFFFFu, small_word, r + 8 * k);
        if (!edge_on[0]) mk[r][0] = 0u;
        if (!edge_on[1]) mk[r][1] = 0u;
        if (!edge_on[2] || mk[r][1] == 0u) mk[r][2] = 0u;
        if ((mk[r][0] | mk[r][1]) == 0u) bulk_mask |= 1u << r;
    }
#ifdef PP_M3_TIMING
    { uint32_t acc = 0;
#pragma unroll
      for (int r = 0; r < RPC; r++) acc |= mk[r][0] | mk[r][1];
      if (acc == 0x12345u) pp_m3_clk[0] = 1; }
    M3_CLK(12);  // masks have arrived
#endif
    // Systems with crossing lanes: the neighbouring slots' words are read with plain loads (two replicas per round, all loads
    // of a round in flight together, issued before the wait for the bulk copies) and merged into the staged words.
    if (bulk_mask != (1u << RPC) - 1u && pt_schedule == 0) {
        // single_random_edge: every (replica, side) with a crossing lane is one exchange  a ^= (a ^ nbr) & mask  (a lane attempts
        // one edge, so the two masks of a replica are disjoint and the exchanges commute).  Up to four exchanges per round, the
        // loads of a round all in flight before the first merge: one round for nearly every CTA.
        const uint32_t n4 = N / 4;
        uint32_t todo = 0;  // bit 2r + side
#pragma unroll
        for (int r = 0; r < RPC; r++) todo |= (mk[r][0] ? 1u << (2 * r) : 0u) | (mk[r][1] ? 2u << (2 * r) : 0u);
        bool waited = false;
        while (todo) {
            int e[4];
            uint32_t em[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                e[j] = todo ? __ffs((int)todo) - 1 : -1;
                todo &= todo - 1u;
                em[j] = 0u;
#pragma unroll
                for (int r = 0; r < RPC; r++) {
                    if (e[j] == 2 * r) em[j] = mk[r][0];
                    if (e[j] == 2 * r + 1) em[j] = mk[r][1];
                }
            }
            if (NFIX == 16 * MSC3D_NTH) {  // four 128-bit words per thread and system: no loop, no bounds
                uint4 nb[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (e[j] < 0) continue;
                    const int r = e[j] >> 1, tn = t + ((e[j] & 1) ? 1 : -1);
                    const uint4 *src = reinterpret_cast<const uint4 *>(m.words + ((g * m.S + (int64_t)r * m.T + tn) * (int64_t)N)) + ht;
#pragma unroll
                    for (int k = 0; k < 4; k++) nb[j][k] = __ldg(src + k * MSC3D_NTH);
                }
                if (!waited) { wait_spins(); waited = true; }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (e[j] < 0) continue;
                    uint4 *dst = reinterpret_cast<uint4 *>(sp + (size_t)(e[j] >> 1) * N) + ht;
                    const uint32_t mm = em[j];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        uint4 a = dst[k * MSC3D_NTH];
                        const uint4 b = nb[j][k];
                        a.x ^= (a.x ^ b.x) & mm; a.y ^= (a.y ^ b.y) & mm; a.z ^= (a.z ^ b.z) & mm; a.w ^= (a.w ^ b.w) & mm;
                        dst[k * MSC3D_NTH] = a;
                    }
                }
            } else for (uint32_t i0 = 0; i0 < n4; i0 += 4 * MSC3D_NTH) {
                uint4 nb[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (e[j] < 0) continue;
                    const int r = e[j] >> 1, tn = t + ((e[j] & 1) ? 1 : -1);
                    const uint4 *src = reinterpret_cast<const uint4 *>(m.words + ((g * m.S + (int64_t)r * m.T + tn) * (int64_t)N));
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * MSC3D_NTH + ht;
                        if (i < n4) nb[j][k] = __ldg(src + i);
                    }
                }
#ifdef PP_M3_TIMING
                if (!waited) M3_CLK(13);  // gather loads issued
#endif
                if (!waited) { wait_spins(); waited = true; M3_CLK(14); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (e[j] < 0) continue;
                    uint4 *dst = reinterpret_cast<uint4 *>(sp + (size_t)(e[j] >> 1) * N);
                    const uint32_t mm = em[j];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * MSC3D_NTH + ht;
                        if (i < n4) {
                            uint4 a = dst[i];
                            const uint4 b = nb[j][k];
                            a.x ^= (a.x ^ b.x) & mm; a.y ^= (a.y ^ b.y) & mm; a.z ^= (a.z ^ b.z) & mm; a.w ^= (a.w ^ b.w) & mm;
                            dst[i] = a;
                        }
                    }
                }
            }
        }
    } else if (bulk_mask != (1u << RPC) - 1u) {
        const uint32_t n4 = N / 4;
        bool waited = false;
#pragma unroll
        for (int rb = 0; rb < RPC; rb += 2) {
            if ((((~bulk_mask) >> rb) & 3u) == 0u) continue;
            for (uint32_t i0 = 0; i0 < n4; i0 += 4 * MSC3D_NTH) {
                uint4 nb[2][4][3];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int r = rb + rr;
                    if (r >= RPC || (bulk_mask >> r & 1u)) continue;
                    const uint32_t *base = m.words + (g * m.S + (int64_t)r * m.T) * (int64_t)N;  // slot 0 of this replica's ladder
                    const uint4 *s0, *s1, *s2;
                    if (pt_schedule == 0) {
                        s0 = reinterpret_cast<const uint4 *>(base + (int64_t)(t - 1) * N);
                        s1 = reinterpret_cast<const uint4 *>(base + (int64_t)(t + 1) * N);
                        s2 = s1;
                    } else {
                        s0 = reinterpret_cast<const uint4 *>(base + (int64_t)max(t1, 0) * N);
                        s1 = reinterpret_cast<const uint4 *>(base + (int64_t)max(t2, 0) * N);
                        s2 = reinterpret_cast<const uint4 *>(base + (int64_t)max(t21, 0) * N);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * MSC3D_NTH + ht;
                        if (i < n4) {
                            if (mk[r][0]) nb[rr][k][0] = __ldg(s0 + i);
                            if (mk[r][1]) nb[rr][k][1] = __ldg(s1 + i);
                            if (pt_schedule == 1 && mk[r][1] && mk[r][2]) nb[rr][k][2] = __ldg(s2 + i);
                        }
                    }
                }
                if (!waited) { wait_spins(); waited = true; }
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int r = rb + rr;
                    if (r >= RPC || (bulk_mask >> r & 1u)) continue;
                    uint4 *dst = reinterpret_cast<uint4 *>(sp + (size_t)r * N);
                    const uint32_t m0 = mk[r][0], m1 = mk[r][1], m2 = mk[r][2];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * MSC3D_NTH + ht;
                        if (i < n4) {
                            uint4 a = dst[i];
                            if (pt_schedule == 0) {
                                uint4 o = a;
                                if (m0) { const uint4 b = nb[rr][k][0]; o.x ^= (a.x ^ b.x) & m0; o.y ^= (a.y ^ b.y) & m0; o.z ^= (a.z ^ b.z) & m0; o.w ^= (a.w ^ b.w) & m0; }
                                if (m1) { const uint4 c = nb[rr][k][1]; o.x ^= (a.x ^ c.x) & m1; o.y ^= (a.y ^ c.y) & m1; o.z ^= (a.z ^ c.z) & m1; o.w ^= (a.w ^ c.w) & m1; }
                                a = o;
                            } else {
                                if (m0) { const uint4 b = nb[rr][k][0]; a.x ^= (a.x ^ b.x) & m0; a.y ^= (a.y ^ b.y) & m0; a.z ^= (a.z ^ b.z) & m0; a.w ^= (a.w ^ b.w) & m0; }
                                if (m1) {
                                    uint4 c = nb[rr][k][1];
                                    if (m2) { const uint4 e = nb[rr][k][2]; c.x ^= (c.x ^ e.x) & m2; c.y ^= (c.y ^ e.y) & m2; c.z ^= (c.z ^ e.z) & m2; c.w ^= (c.w ^ e.w) & m2; }
                                    a.x ^= (a.x ^ c.x) & m1; a.y ^= (a.y ^ c.y) & m1; a.z ^= (a.z ^ c.z) & m1; a.w ^= (a.w ^ c.w) & m1;
                                }
                            }
                            dst[i] = a;
                        }
                    }
                }
            }
        }
    }
    M3_CLK(1);
    if (ids_pre) {
#pragma unroll
        for (int r = 0; r < (RPC <= 4 ? RPC : 1); r++) ids[r * 32 + ht] = (uint16_t)my_ids[r];
    }
    uint64_t nthr[3];
#pragma unroll
    for (int u = 0; u < 3; u++)  // (one - 1) * tid = 0: keeps the addends in per-thread registers, where IMAD.WIDE can take them
        nthr[u] = 0ull - (uint64_t)thr[u] + (uint64_t)((gv.one[u] - 1u) * (uint32_t)tid);
    const uint64_t key = msc_group_key(m.seed, (uint64_t)(group_offset + g));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    mbar_wait(&bars[0], 0);
    wait_spins();
    if (bulk_mask != (1u << RPC) - 1u) half_barrier(half, MSC3D_NTH);  // gathered words were written with plain stores
    M3_CLK(2);

    // ---- sweeps: colour 0 then colour 1 (RNG-SPEC visit order)
    // In-sweep energy (esw, host-checked: one slot per CTA, an even number of replicas, at most two quads per thread and
    // room for the counters in the coupling-word area): the colour-1 pass of the last sweep also counts the unsatisfied
    // bonds, and the epilogue below only has the replica pairs left to walk.
    constexpr bool ESW_OK = NH == 1 && RPC >= 2;
    const bool esw_on = ESW_OK && esw && n_sweeps > 0 && want_energy && (want_mags != 0) == (want_overlap != 0);
    VAcc<MSC3D_KS> ea[RPC];
#pragma unroll
    for (int r = 0; r < RPC; r++) ea[r].clear();
    for (int sw = 0; sw < n_sweeps; sw++) {
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
            const uint32_t so = c * N2, oo = (1 - c) * N2;  // word offsets of the updated / the other colour half
            const uint32_t tag = TAG_SWEEP_MSC | (uint32_t)c;
            uint4 next = ht < gv.n_items ? item_at(ht) : make_uint4(0, 0, 0, 0);
            if (ESW_OK && esw_on && c == 1 && sw == n_sweeps - 1) {
#pragma unroll 1
                for (int it = ht; it < gv.n_items; it += MSC3D_NTH) {
                    const uint4 desc = next;
                    if (it + MSC3D_NTH < gv.n_items) next = item_at(it + MSC3D_NTH);
                    msc3d_sweep_item<RPC, METRO, ESW_OK>(sp, J, N, so, oo, desc, (((desc.w >> 16) ^ 1u) & 1u) != 0u, thr, nthr, gv.one,
                                                         sweep_index + (uint32_t)sw, (uint32_t)t, (uint32_t)m.T, tag, k0, k1, ea);
                }
            } else {
#pragma unroll 1
                for (int it = ht; it < gv.n_items; it += MSC3D_NTH) {
                    const uint4 desc = next;
                    if (it + MSC3D_NTH < gv.n_items) next = item_at(it + MSC3D_NTH);
                    msc3d_sweep_item<RPC, METRO, false>(sp, J, N, so, oo, desc, (((desc.w >> 16) ^ (uint32_t)c) & 1u) != 0u, thr, nthr,
                                                        gv.one, sweep_index + (uint32_t)sw, (uint32_t)t, (uint32_t)m.T, tag, k0, k1, nullptr);
                }
            }
            half_barrier(half, MSC3D_NTH);
            M3_CLK(3 + c);
        }
    }

    // ---- stage out (asynchronous; the epilogue below only reads shared memory)
    if (n_sweeps > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        half_barrier(half, MSC3D_NTH);
        if (ht == 0) {
#pragma unroll
            for (int r = 0; r < RPC; r++)
                bulk_s2g(words_out + ((g * m.S + (int64_t)r * m.T + t) * (int64_t)N), sp + (size_t)r * N, bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }

    M3_CLK(5);
    // ---- epilogue: per-lane unsatisfied forward bonds, down spins, replica-pair overlaps.
    // Warps 0..3 of the half: energy + magnetisation, 4/RPC warps per replica; warps 4..7: the P replica pairs,
    // 4/P warps per pair.  Each warp strides over the row segments, accumulates bit-sliced counters and reduces them
    // with a bit-sliced butterfly; partial lane totals meet in shared memory.
    if (want_energy || want_mags || want_overlap) {
        constexpr int NP = RPC / 2;              // replica pairs
        const int w = ht >> 5, lane = ht & 31;
        // where the final lane totals end up: [replica r][E, M][fin_wpe][32] and [pair p][q, ql][fin_wpp][32]
        const uint32_t *fin_e = nullptr, *fin_p = nullptr;
        int fin_wpe = 1, fin_wpp = 1;
        if (ESW_OK && esw_on) {
            // ---- in-sweep energy path.  The coupling words are dead after the last pass: their area takes the per-thread
            // counters (merged across the eight warps with bit-sliced adds, one butterfly per quantity) and the lane totals.
            constexpr int WPPX = NP > 0 ? 8 / NP : 8;  // warps per pair: all eight warps walk the pairs
            uint32_t *scr = smem;
            uint32_t *res = smem + 3 * N - 512;
            uint32_t *res_p = res + RPC * 2 * 32;
#pragma unroll
            for (int r = 0; r < RPC; r++)
#pragma unroll
                for (int b = 0; b < MSC3D_KS; b++) scr[(r * MSC3D_KS + b) * MSC3D_NTH + ht] = ea[r].p[b];
            M3_T(6);
            half_barrier(half, MSC3D_NTH);
            if (w < RPC) res[(w * 2 + 0) * 32 + lane] = merge_lane_total<MSC3D_KS, 8>(scr + w * MSC3D_KS * MSC3D_NTH, lane);
            M3_T(7);
            if (NP > 0 && want_overlap) {
                const int p = w / WPPX, sub = w % WPPX;
                const uint32_t *A = sp + (size_t)(2 * p) * N, *B = sp + (size_t)(2 * p + 1) * N;
                VAcc<MSC3D_KL> vl;
                VAcc<MSC3D_KQ> vq, vma, vmb;
                vl.clear(); vq.clear(); vma.clear(); vmb.clear();
                const int it0 = lane + 32 * sub;
                uint4 next = it0 < gv.n_items ? item_at(it0) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
                for (int it = it0; it < gv.n_items; it += 32 * WPPX) {
                    const uint4 desc = next;
                    if (it + 32 * WPPX < gv.n_items) next = item_at(it + 32 * WPPX);
                    msc3d_pairm_item(A, B, N2, desc, vl, vq, vma, vmb);
                }
                M3_T(8);
                half_barrier(half, MSC3D_NTH);  // the energy counters have been consumed
                // park: pair p at scr + p * PW, quantities [q | ql | M_a | M_b], each [plane][WPPX * 32]
                constexpr int QW = MSC3D_KQ * WPPX * 32, LW = MSC3D_KL * WPPX * 32, PW = 3 * QW + LW;
                uint32_t *base = scr + p * PW + sub * 32 + lane;
#pragma unroll
                for (int b = 0; b < MSC3D_KQ; b++) {
                    base[b * WPPX * 32] = vq.p[b];
                    base[QW + LW + b * WPPX * 32] = vma.p[b];
                    base[2 * QW + LW + b * WPPX * 32] = vmb.p[b];
                }
#pragma unroll
                for (int b = 0; b < MSC3D_KL; b++) base[QW + b * WPPX * 32] = vl.p[b];
                half_barrier(half, MSC3D_NTH);
                M3_T(9);
                if (w < NP * 4) {  // one warp per (pair, quantity)
                    const int p2 = w >> 2, qn = w & 3;
                    const uint32_t *src = scr + p2 * PW;
                    // qn: 0 = q, 1 = q_link, 2 = down spins of replica 2p, 3 = of replica 2p + 1
                    const uint32_t *q_src = src + (qn == 0 ? 0 : qn == 1 ? QW : qn == 2 ? QW + LW : 2 * QW + LW);
                    uint32_t *q_dst = qn < 2 ? res_p + (p2 * 2 + qn) * 32 : res + ((2 * p2 + qn - 2) * 2 + 1) * 32;
                    q_dst[lane] = qn == 1 ? merge_lane_total<MSC3D_KL, WPPX>(q_src, lane) : merge_lane_total<MSC3D_KQ, WPPX>(q_src, lane);
                }
            }
            half_barrier(half, MSC3D_NTH);
            M3_T(10);
            fin_e = res; fin_p = res_p;
        } else {
        constexpr int WPP = NP > 0 ? 4 / NP : 1; // warps per pair
        // energy / magnetisation: warps 0..3 when the other four take the replica pairs, else all eight
        const bool pairs_on = NP > 0 && want_overlap;
        const int n_em = pairs_on ? 4 : 8, WPE = n_em / RPC;  // warps per replica
        // red layout: [replica r][E, M][WPE][32] then [pair p][q, ql][WPP][32]
        uint32_t *red_p = red + RPC * 2 * WPE * 32;
        uint32_t tot0 = 0, tot1 = 0;  // this warp's two lane totals: (E, M) or (q, ql)
        uint32_t *slot0 = nullptr, *slot1 = nullptr;
        if (w < n_em) {
            if (want_energy || want_mags) {
                const int r = w / WPE, sub = w % WPE;
                const uint32_t *A = sp + (size_t)r * N;
                VAcc<MSC3D_KE> ve;
                VAcc<MSC3D_KM> vm;
                ve.clear(); vm.clear();
#pragma unroll 1
                for (int it = lane + 32 * sub; it < gv.n_items; it += 32 * WPE) {
                    const uint4 desc = item_at(it);
                    msc3d_em_item(A, J, N, N2, desc, want_energy, want_mags, ve, vm);
                }
                if (want_energy) { tot0 = warp_lane_total(ve); slot0 = red + ((r * 2 + 0) * WPE + sub) * 32 + lane; }
                if (want_mags) { tot1 = warp_lane_total(vm); slot1 = red + ((r * 2 + 1) * WPE + sub) * 32 + lane; }
            }
        } else if (NP > 0 && want_overlap) {
            const int pw = w - 4, p = pw / WPP, sub = pw % WPP;
            if (p < m.P) {
                const uint32_t *A = sp + (size_t)(2 * p) * N, *B = sp + (size_t)(2 * p + 1) * N;
                VAcc<MSC3D_KE> vl;
                VAcc<MSC3D_KM> vq;
                vl.clear(); vq.clear();
#pragma unroll 1
                for (int it = lane + 32 * sub; it < gv.n_items; it += 32 * WPP) {
                    const uint4 desc = item_at(it);
                    msc3d_pair_item(A, B, N2, desc, vl, vq);
                }
                tot0 = warp_lane_total(vq); slot0 = red_p + ((p * 2 + 0) * WPP + sub) * 32 + lane;
                tot1 = warp_lane_total(vl); slot1 = red_p + ((p * 2 + 1) * WPP + sub) * 32 + lane;
            }
        }
        if (NH == 1) {  // the scratch aliases the spin buffer: every warp must be done reading it, and so must the bulk store
            if (n_sweeps > 0 && ht == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            half_barrier(half, MSC3D_NTH);
        }
        if (slot0) *slot0 = tot0;
        if (slot1) *slot1 = tot1;
        half_barrier(half, MSC3D_NTH);
        fin_e = red; fin_p = red_p; fin_wpe = WPE; fin_wpp = WPP;
        }
        // Lane l finishes realization 32g + l at slot t -- per-system energy / magnetisation, pair dots, and (want_fold) the
        // recorded-sweep fold of simulation/mod.rs:543-578 + statistics/overlap.rs:283-306.  A "unit" is a replica (u < RPC) or a
        // replica pair (u - RPC).  unit_values: lane totals -> f32 observables + the per-system / per-pair outputs; unit_fold: the
        // f64 reductions, same values and same order as fold_red (pp_kernels_stats.cuh): replica-major, then pair-major.
        // One slot per CTA: the units' values are worked out by one warp each (they are ~300 dependent instructions apiece),
        // parked, and warp 0 only issues the ordered reductions; otherwise warp 0 walks the units.
        // Rolled on purpose: unrolled (with inlined f32 divisions) this was ~1200 instructions of straight-line code
        // streaming through the instruction cache the hot loops need.
        const int64_t d = g * 32 + lane;
        const float nf = (float)m.N, nb = (float)(m.N * m.z);
        const int n_units = RPC + ((NP > 0 && want_overlap) ? NP : 0);
        auto unit_values = [&](const int u, float &v0, float &v1, int &v2) {
            if (u < RPC) {
                const int r = u;
                uint32_t e = 0, dn = 0;
                for (int k = 0; k < fin_wpe; k++) {
                    if (want_energy) e += fin_e[((r * 2 + 0) * fin_wpe + k) * 32 + lane];
                    if (want_mags) dn += fin_e[((r * 2 + 1) * fin_wpe + k) * 32 + lane];
                }
                // sum_i sum_d s s J = (#bonds) - 2 * unsatisfied, e = that / N in f32   (energy.rs:103-108)
                const float ev = __fdiv_rn((float)(3ll * N - 2ll * e), (float)N);
                const long long mv = (long long)N - 2ll * dn;
                const int sys = ids_ok ? (int)ids[r * 32 + lane] : m.system_ids[d * m.S + r * m.T + t];
                if (want_energy) m.energies[d * m.S + sys] = ev;
                if (want_mags) m.mags[d * m.S + sys] = mv;
                v0 = ev;
                v1 = want_fold ? __fdiv_rn((float)mv, nf) : 0.0f;
                v2 = 0;
            } else {
                const int p = u - RPC;
                uint32_t cs = 0, cl = 0;
                for (int k = 0; k < fin_wpp; k++) {
                    cs += fin_p[((p * 2 + 0) * fin_wpp + k) * 32 + lane];
                    cl += fin_p[((p * 2 + 1) * fin_wpp + k) * 32 + lane];
                }
                const long long sv = (long long)N - 2ll * cs, lv = 3ll * N - 2ll * cl;
                const int64_t o = (d * m.P + p) * m.T + t;
                dot_spin[o] = sv;
                dot_link[o] = lv;
                v0 = want_fold ? __fdiv_rn((float)sv, nf) : 0.0f;
                v1 = want_fold ? __fdiv_rn((float)lv, nb) : 0.0f;
                v2 = (int)sv;
            }
        };
        auto unit_fold = [&](const int u, const float v0, const float v1, const int v2) {
            double *sums = st.sums + d * 11 * m.T + t;
            if (u < RPC) {  // simulation/mod.rs:555-578
                const float ev = v0, mag = v1;
                const float m2 = __fmul_rn(mag, mag);
                atomicAdd(sums + 0 * m.T, (double)mag);
                atomicAdd(sums + 1 * m.T, (double)m2);
                atomicAdd(sums + 2 * m.T, (double)__fmul_rn(m2, m2));
                atomicAdd(sums + 3 * m.T, (double)ev);
                atomicAdd(sums + 4 * m.T, __dmul_rn((double)ev, (double)ev));
            } else {  // statistics/overlap.rs:283-306 + :318-324
                const float q = v0, ql = v1;
                const float q2 = __fmul_rn(q, q);
                const float ql2 = __fmul_rn(ql, ql);
                atomicAdd(sums + 5 * m.T, (double)q);
                atomicAdd(sums + 6 * m.T, (double)q2);
                atomicAdd(sums + 7 * m.T, (double)__fmul_rn(q2, q2));
                atomicAdd(sums + 8 * m.T, (double)ql);
                atomicAdd(sums + 9 * m.T, (double)ql2);
                atomicAdd(sums + 10 * m.T, (double)__fmul_rn(ql2, ql2));
                const int64_t h = (d * m.T + t) * (int64_t)(m.N + 1) + ((long long)v2 + m.N) / 2;
                atomicAdd(st.hist + h, 1u);
                atomicAdd(st.ql_at_q + h, (double)ql);
                atomicAdd(st.ql2_at_q + h, (double)ql2);
            }
        };
        if (NH == 1 && RPC <= 4) {
            // parked values [unit][3][32]: the (dead) coupling words on the in-sweep-energy path, else behind the lane totals
            float *park = reinterpret_cast<float *>((ESW_OK && esw_on) ? smem : red + 512);
            if (d < m.D) {
#pragma unroll 1
                for (int u = w; u < n_units; u += 8) {
                    float v0, v1;
                    int v2;
                    unit_values(u, v0, v1, v2);
                    park[(u * 3 + 0) * 32 + lane] = v0;
                    park[(u * 3 + 1) * 32 + lane] = v1;
                    park[(u * 3 + 2) * 32 + lane] = __int_as_float(v2);
                }
            }
            if (want_fold) {
                M3_T(11);
                half_barrier(half, MSC3D_NTH);
                if (w == 0 && d < m.D) {
#pragma unroll 1
                    for (int u = 0; u < n_units; u++)
                        unit_fold(u, park[(u * 3 + 0) * 32 + lane], park[(u * 3 + 1) * 32 + lane], __float_as_int(park[(u * 3 + 2) * 32 + lane]));
                }
            }
        } else if (w == 0 && d < m.D) {
#pragma unroll 1
            for (int u = 0; u < n_units; u++) {
                float v0, v1;
                int v2;
                unit_values(u, v0, v1, v2);
                if (want_fold) unit_fold(u, v0, v1, v2);
            }
        }
    }
    M3_CLK(6);
    if (n_sweeps > 0 && ht == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    M3_CLK(7);
#ifdef PP_M3_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 8192) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); pp_m3_clk[blockIdx.x * 32 + 9] = gt; }
#endif
}
#endif  // __CUDACC__

}  // namespace pp
