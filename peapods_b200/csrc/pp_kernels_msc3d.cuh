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
// bulk-async store per system = exactly the algorithmic 2 x 4 B per word), the sign words of the
// couplings are read through L1 once per quad and reused by the R replicas.
//
// Per word update (Metropolis): 6 three-input XORs (bond words), two bit-sliced full adders and five
// LOP3s give the masks [unsat >= 1], [>= 2], [>= 3]; the 24-bit draw (one Philox4x32-10 call per quad
// and replica, RNG-SPEC TAG_SWEEP_MSC) selects which of them is the flip mask:
//   flip lane l  <=>  draw < table[t][2*unsat_l]   (sweep.rs:182-184 with ec + 2z' = 2*unsat).
#pragma once
#include <string>
#include <vector>

#include "../../include/peapods_b200.h"
#include "pp_device.cuh"
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
    for (int it = 0; it < q.n_items; it++) {
        const int row = it / q.QPR, h = it % q.QPR, x0 = row / q.L1, x1 = row % q.L1;
        auto at = [&](int a, int b, int j) { return (uint16_t)((a * q.L1 + b) * q.LXH + j); };
        uint16_t *d = &q.items[(size_t)it * 8];
        d[0] = at(x0, x1, 4 * h);
        d[1] = at((x0 + 1) % q.L0, x1, 4 * h);
        d[2] = at((x0 + q.L0 - 1) % q.L0, x1, 4 * h);
        d[3] = at(x0, (x1 + 1) % q.L1, 4 * h);
        d[4] = at(x0, (x1 + q.L1 - 1) % q.L1, 4 * h);
        d[5] = at(x0, x1, (4 * h + q.LXH - 1) % q.LXH);
        d[6] = at(x0, x1, (4 * h + 4) % q.LXH);
        d[7] = (uint16_t)((x0 + x1) & 1);
    }
    q.ok = true;
    return q;
}

struct Msc3dView {
    const uint4 *items;  // [n_items] packed 8 x u16
    int n_items;
    uint32_t N, N2;
};

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    // one word into plane LVL, return the carry
    template <int LVL>
    __device__ __forceinline__ uint32_t half(uint32_t a) {
        const uint32_t c = p[LVL] & a;
        p[LVL] ^= a;
        return c;
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

// Sum the K-plane counters of the 32 threads of a warp (bit-sliced butterfly), then thread l returns
// the total of lane l.
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

__device__ __forceinline__ uint4 lds4(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint4 ldg4(const uint32_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

constexpr int MSC3D_KE = 10;  // planes of the per-thread unsatisfied-bond counter   (3*sites_per_thread < 1024)
constexpr int MSC3D_KM = 9;   // planes of the per-thread down-spin / q counters     (sites_per_thread   < 512)

// ------------------------------------------------------------------------------------------------
// grid.x = G*T (word group, temperature slot); block = NT threads, NT % (32*RPC) == 0.
// dynamic smem: RPC*N words + 16 B (mbarrier) + RPC*4*(NT/RPC/32)*32 words (reduction scratch)
template <int RPC, bool METRO>
__global__ void __launch_bounds__(256, 2)
msc3d_kernel(ModelView m, Msc3dView gv, uint32_t sweep_index, int n_sweeps, int want_energy, int want_mags,
             int want_overlap, int64_t group_offset, long long *dot_spin, long long *dot_link) {
    extern __shared__ __align__(128) uint32_t smem[];
    const uint32_t N = gv.N, N2 = gv.N2;
    uint32_t *sp = smem;                                                    // [RPC][N]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + (size_t)RPC * N);
    uint32_t *red = smem + (size_t)RPC * N + 4;                             // reduction scratch
    const int tid = threadIdx.x, NT = blockDim.x;
    const int64_t g = blockIdx.x / m.T;
    const int t = blockIdx.x % m.T;
    const uint32_t bytes = N * 4u;

    // ---- stage in: one bulk-async copy per system, completion on an mbarrier
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes * RPC) : "memory");
#pragma unroll
        for (int r = 0; r < RPC; r++) {
            const uint32_t *src = m.words + ((g * m.S + (int64_t)r * m.T + t) * (int64_t)N);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(sp + (size_t)r * N)),
                         "l"(src), "r"(bytes), "r"(smem_u32(bar))
                         : "memory");
        }
    }
    // acceptance counts of this temperature: cnt[u] = table[t][2u]  (sweep.rs:162-166, index ec + 2z' = 2*unsat)
    uint32_t cnt[7];
#pragma unroll
    for (int u = 0; u < 7; u++) cnt[u] = m.lut[t * 13 + 2 * u];
    const uint64_t key = msc_group_key(m.seed, (uint64_t)(group_offset + g));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint32_t *Jg = m.Jw ? m.Jw + g * 3 * (int64_t)N : nullptr;
    __syncthreads();  // mbarrier init visible to all waiters
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(smem_u32(bar))
                : "memory");
        }
    }

    // ---- sweeps: colour 0 then colour 1 (RNG-SPEC visit order)
    for (int sw = 0; sw < n_sweeps; sw++) {
#pragma unroll 1
        for (int c = 0; c < 2; c++) {
            const uint32_t so = c * N2, oo = (1 - c) * N2;  // word offsets of the updated / the other colour half
#pragma unroll 1
            for (int it = tid; it < gv.n_items; it += NT) {
                uint32_t self, zp, zm, yp, ym, eL, eR, par;
                unpack_item(__ldg(gv.items + it), self, zp, zm, yp, ym, eL, eR, par);
                const bool p = ((par ^ (uint32_t)c) & 1u) != 0;  // sites of this colour in the row sit at x2 = 2j + p
                // coupling sign words (bit = 1: J = -1); bond (i, d) is stored at its lower site i
                uint32_t Jf0[4], Jf1[4], Jf2[4], Jb0[4], Jb1[4], Jb2[4];
                if (Jg) {
                    uint4 v;
                    v = ldg4(Jg + 0 * N + so + self); Jf0[0] = v.x; Jf0[1] = v.y; Jf0[2] = v.z; Jf0[3] = v.w;
                    v = ldg4(Jg + 1 * N + so + self); Jf1[0] = v.x; Jf1[1] = v.y; Jf1[2] = v.z; Jf1[3] = v.w;
                    v = ldg4(Jg + 2 * N + so + self); Jf2[0] = v.x; Jf2[1] = v.y; Jf2[2] = v.z; Jf2[3] = v.w;
                    v = ldg4(Jg + 0 * N + oo + zm);   Jb0[0] = v.x; Jb0[1] = v.y; Jb0[2] = v.z; Jb0[3] = v.w;
                    v = ldg4(Jg + 1 * N + oo + ym);   Jb1[0] = v.x; Jb1[1] = v.y; Jb1[2] = v.z; Jb1[3] = v.w;
                    v = ldg4(Jg + 2 * N + oo + self);
                    if (p) { Jb2[0] = v.x; Jb2[1] = v.y; Jb2[2] = v.z; Jb2[3] = v.w; }
                    else   { Jb2[0] = __ldg(Jg + 2 * N + oo + eL); Jb2[1] = v.x; Jb2[2] = v.y; Jb2[3] = v.z; }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) Jf0[j] = Jf1[j] = Jf2[j] = Jb0[j] = Jb1[j] = Jb2[j] = 0u;
                }
#pragma unroll
                for (int r = 0; r < RPC; r++) {
                    uint32_t *sys = sp + (size_t)r * N;
                    const u32x4 rnd = philox4x32_10(self >> 2, sweep_index + (uint32_t)sw, (uint32_t)(r * m.T + t),
                                                    TAG_SWEEP_MSC | (uint32_t)c, k0, k1);
                    const uint4 S = lds4(sys + so + self);
                    const uint4 ZP = lds4(sys + oo + zp), ZM = lds4(sys + oo + zm);
                    const uint4 YP = lds4(sys + oo + yp), YM = lds4(sys + oo + ym);
                    const uint4 O = lds4(sys + oo + self);
                    const uint32_t E = sys[oo + (p ? eR : eL)];
                    uint32_t s[4] = {S.x, S.y, S.z, S.w};
                    const uint32_t zpv[4] = {ZP.x, ZP.y, ZP.z, ZP.w}, zmv[4] = {ZM.x, ZM.y, ZM.z, ZM.w};
                    const uint32_t ypv[4] = {YP.x, YP.y, YP.z, YP.w}, ymv[4] = {YM.x, YM.y, YM.z, YM.w};
                    uint32_t xr[4], xl[4];
                    if (p) { xl[0] = O.x; xl[1] = O.y; xl[2] = O.z; xl[3] = O.w; xr[0] = O.y; xr[1] = O.z; xr[2] = O.w; xr[3] = E; }
                    else   { xl[0] = E;   xl[1] = O.x; xl[2] = O.y; xl[3] = O.z; xr[0] = O.x; xr[1] = O.y; xr[2] = O.z; xr[3] = O.w; }
                    const uint32_t draw[4] = {rnd.x >> 8, rnd.y >> 8, rnd.z >> 8, rnd.w >> 8};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b0 = xor3(s[j], zpv[j], Jf0[j]), b1 = xor3(s[j], zmv[j], Jb0[j]);
                        const uint32_t b2 = xor3(s[j], ypv[j], Jf1[j]), b3 = xor3(s[j], ymv[j], Jb1[j]);
                        const uint32_t b4 = xor3(s[j], xr[j], Jf2[j]), b5 = xor3(s[j], xl[j], Jb2[j]);
                        const uint32_t s1 = xor3(b0, b1, b2), c1 = maj3(b0, b1, b2);
                        const uint32_t s2 = xor3(b3, b4, b5), c2 = maj3(b3, b4, b5);
                        const uint32_t kk = s1 & s2, oo2 = s1 | s2;
                        uint32_t flip;
                        if (METRO) {  // counts for unsat >= 3 are 2^24 (ec >= 0 always accepts, sweep.rs:141-145)
                            const uint32_t ge1 = oo2 | c1 | c2, ge2 = kk | c1 | c2, ge3 = maj3(c1, c2, oo2);
                            flip = ge3;
                            if (draw[j] < cnt[2]) flip = ge2;
                            if (draw[j] < cnt[1]) flip = ge1;
                            if (draw[j] < cnt[0]) flip = 0xFFFFFFFFu;
                        } else {
                            const uint32_t x0 = s1 ^ s2, y1 = xor3(c1, c2, kk), y2 = maj3(c1, c2, kk);
                            flip = 0u;
#pragma unroll
                            for (int u = 0; u < 7; u++) {
                                const uint32_t eq = ((u & 1) ? x0 : ~x0) & ((u & 2) ? y1 : ~y1) & ((u & 4) ? y2 : ~y2);
                                if (draw[j] < cnt[u]) flip |= eq;
                            }
                        }
                        s[j] ^= flip;
                    }
                    *reinterpret_cast<uint4 *>(sys + so + self) = make_uint4(s[0], s[1], s[2], s[3]);
                }
            }
            __syncthreads();
        }
    }

    // ---- stage out (asynchronous; the epilogue below only reads shared memory)
    if (n_sweeps > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < RPC; r++) {
                uint32_t *dst = m.words + ((g * m.S + (int64_t)r * m.T + t) * (int64_t)N);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                             "r"(smem_u32(sp + (size_t)r * N)), "r"(bytes)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }

    // ---- epilogue: per-lane unsatisfied forward bonds, down spins, replica-pair overlaps
    if (want_energy || want_mags || want_overlap) {
        const int TPR = NT / RPC;          // threads per replica (multiple of 32)
        const int r = tid / TPR, jt = tid % TPR;
        const int wpr = TPR >> 5, wr = jt >> 5, lane = tid & 31;
        const uint32_t *A = sp + (size_t)r * N;
        const uint32_t *B = sp + (size_t)(r ^ 1) * N;
        const bool paired = want_overlap && (r ^ 1) < RPC && (r | 1) < 2 * m.P;
        VAcc<MSC3D_KE> ve, vl;
        VAcc<MSC3D_KM> vm, vq;
        ve.clear(); vl.clear(); vm.clear(); vq.clear();
        int iter = 0;
#pragma unroll 1
        for (int it = jt; it < gv.n_items; it += TPR, iter++) {
            uint32_t self, zp, zm, yp, ym, eL, eR, par;
            unpack_item(__ldg(gv.items + it), self, zp, zm, yp, ym, eL, eR, par);
            (void)zm; (void)ym; (void)eL;
            // the 8 sites of the segment: words 0..3 colour 0, 4..7 colour 1
            uint32_t a[8], az[8], ay[8], ax[8];
            {
                const uint4 s0 = lds4(A + self), s1 = lds4(A + N2 + self);
                const uint4 z0 = lds4(A + N2 + zp), z1 = lds4(A + zp);    // +x0 neighbours of colour-0 / colour-1 sites
                const uint4 y0 = lds4(A + N2 + yp), y1 = lds4(A + yp);
                a[0] = s0.x; a[1] = s0.y; a[2] = s0.z; a[3] = s0.w; a[4] = s1.x; a[5] = s1.y; a[6] = s1.z; a[7] = s1.w;
                az[0] = z0.x; az[1] = z0.y; az[2] = z0.z; az[3] = z0.w; az[4] = z1.x; az[5] = z1.y; az[6] = z1.z; az[7] = z1.w;
                ay[0] = y0.x; ay[1] = y0.y; ay[2] = y0.z; ay[3] = y0.w; ay[4] = y1.x; ay[5] = y1.y; ay[6] = y1.z; ay[7] = y1.w;
                // +x2 neighbour: colour c sites sit at x2 = 2j + (par^c); the neighbour 2j + (par^c) + 1 is word j of
                // the other colour when par^c == 0, word j+1 when par^c == 1
                if (par == 0) {  // colour 0: p = 0, colour 1: p = 1
                    ax[0] = s1.x; ax[1] = s1.y; ax[2] = s1.z; ax[3] = s1.w;
                    ax[4] = s0.y; ax[5] = s0.z; ax[6] = s0.w; ax[7] = A[eR];
                } else {         // colour 0: p = 1, colour 1: p = 0
                    ax[0] = s1.y; ax[1] = s1.z; ax[2] = s1.w; ax[3] = A[N2 + eR];
                    ax[4] = s0.x; ax[5] = s0.y; ax[6] = s0.z; ax[7] = s0.w;
                }
            }
            if (want_mags) vm.add8(a);
            if (want_energy) {
                uint32_t j0[8], j1[8], j2[8];
                if (Jg) {
                    uint4 v;
                    v = ldg4(Jg + 0 * N + self);      j0[0] = v.x; j0[1] = v.y; j0[2] = v.z; j0[3] = v.w;
                    v = ldg4(Jg + 0 * N + N2 + self); j0[4] = v.x; j0[5] = v.y; j0[6] = v.z; j0[7] = v.w;
                    v = ldg4(Jg + 1 * N + self);      j1[0] = v.x; j1[1] = v.y; j1[2] = v.z; j1[3] = v.w;
                    v = ldg4(Jg + 1 * N + N2 + self); j1[4] = v.x; j1[5] = v.y; j1[6] = v.z; j1[7] = v.w;
                    v = ldg4(Jg + 2 * N + self);      j2[0] = v.x; j2[1] = v.y; j2[2] = v.z; j2[3] = v.w;
                    v = ldg4(Jg + 2 * N + N2 + self); j2[4] = v.x; j2[5] = v.y; j2[6] = v.z; j2[7] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j++) j0[j] = j1[j] = j2[j] = 0u;
                }
                uint32_t sb[8], cb[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {  // energy.rs:99-107: forward bonds only, each bond once
                    const uint32_t b0 = xor3(a[j], az[j], j0[j]), b1 = xor3(a[j], ay[j], j1[j]), b2 = xor3(a[j], ax[j], j2[j]);
                    sb[j] = xor3(b0, b1, b2);
                    cb[j] = maj3(b0, b1, b2);
                }
                ve.add8_8(sb, cb);
            }
            if (paired && ((iter & 1) == (r & 1))) {  // the two replicas of a pair split the items between them
                uint32_t x[8], sb[8], cb[8];
                const uint4 s0 = lds4(B + self), s1 = lds4(B + N2 + self);
                const uint4 z0 = lds4(B + N2 + zp), z1 = lds4(B + zp);
                const uint4 y0 = lds4(B + N2 + yp), y1 = lds4(B + yp);
                const uint32_t b[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const uint32_t bz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                const uint32_t by[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                uint32_t bx[8];
                if (par == 0) {
                    bx[0] = s1.x; bx[1] = s1.y; bx[2] = s1.z; bx[3] = s1.w;
                    bx[4] = s0.y; bx[5] = s0.z; bx[6] = s0.w; bx[7] = B[eR];
                } else {
                    bx[0] = s1.y; bx[1] = s1.z; bx[2] = s1.w; bx[3] = B[N2 + eR];
                    bx[4] = s0.x; bx[5] = s0.y; bx[6] = s0.z; bx[7] = s0.w;
                }
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
        }
        // cross-thread: warp butterfly, then per-replica combine through shared memory
        uint32_t *mine = red + ((size_t)(r * 4) * wpr + wr) * 32;  // [r][quantity][warp][lane]
        if (want_energy) mine[(size_t)0 * wpr * 32 + lane] = warp_lane_total(ve);
        if (want_mags) mine[(size_t)1 * wpr * 32 + lane] = warp_lane_total(vm);
        if (want_overlap) {
            mine[(size_t)2 * wpr * 32 + lane] = paired ? warp_lane_total(vq) : 0u;
            mine[(size_t)3 * wpr * 32 + lane] = paired ? warp_lane_total(vl) : 0u;
        }
        __syncthreads();
        if (wr == 0) {
            const int64_t d = g * 32 + lane;
            if (d < m.D) {
                const int pos = r * m.T + t;
                auto total = [&](int rr, int q) {
                    uint32_t acc = 0;
                    for (int w = 0; w < wpr; w++) acc += red[(((size_t)rr * 4 + q) * wpr + w) * 32 + lane];
                    return acc;
                };
                const int sys = m.system_ids[d * m.S + pos];
                if (want_energy) {  // sum_i sum_d s s J = (#bonds) - 2 * unsatisfied   (energy.rs:103-108)
                    const long long e_int = 3ll * N - 2ll * total(r, 0);
                    m.energies[d * m.S + sys] = __fdiv_rn((float)e_int, (float)N);
                }
                if (want_mags) m.mags[d * m.S + sys] = (long long)N - 2ll * total(r, 1);
                if (want_overlap && (r & 1) == 0 && r + 1 < 2 * m.P) {
                    const int pr = r >> 1;
                    const int64_t o = (d * m.P + pr) * m.T + t;
                    dot_spin[o] = (long long)N - 2ll * (total(r, 2) + total(r + 1, 2));
                    dot_link[o] = 3ll * N - 2ll * (total(r, 3) + total(r + 1, 3));
                }
            }
        }
    }
    if (n_sweeps > 0 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif  // __CUDACC__

}  // namespace pp
