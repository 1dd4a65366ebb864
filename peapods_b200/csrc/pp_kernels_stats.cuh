// pp_kernels_stats.cuh — scalar bookkeeping kernels shared by both spin layouts.
//
// K8 pt_exchange   <- spin-sim/src/mcmc/tempering.rs:20-102 + PtState simulation/realization.rs:73-120
// K9 fold          <- spin-sim/src/simulation/mod.rs:543-578 + statistics/stats.rs:17-27
//    + K7 overlap moments / histogram scatter <- statistics/overlap.rs:283-306, 318-324
// reduce_hist      <- statistics/overlap.rs:128-142 (sum over realizations, realization order)
//
// One thread per (realization, temperature) or (realization, replica) walks its items
// sequentially so that every f64 accumulation happens in the reference's order.
#pragma once
#include "pp_device.cuh"

namespace pp {

struct StatsView {
    double *sums;          // [D][11][T]: mags, mags2, mags4, energies, energies2, q, q2, q4, ql, ql2, ql4
    uint32_t *hist;        // [D][T][N+1]
    double *ql_at_q;       // [D][T][N+1]
    double *ql2_at_q;      // [D][T][N+1]
    const long long *dot_spin, *dot_link;  // [D][P][T]
};

struct PtView {
    unsigned long long *edge_attempts, *edge_acceptances;  // [D][T-1]
    unsigned long long *round_trips;                       // [D][S] by system
    uint8_t *trip_state;                                   // [D][S] by system
    uint32_t *swap_mask;   // MSC only: [G][R][T-1] lanes whose configurations swap across that edge
    int cold_slot, hot_slot;
};

// K9 (+K7): fold one recorded sweep for (realization d, temperature t).
// Mv / Ev: magnetisation sums and energies of the R systems sitting at slot t (replica order);
// dsp / dlk: the P pair dots at slot t.  All f64 accumulations in the reference's order.
// RFIX > 0: the replica count is the compile-time constant RFIX (loops unroll, the getters may index registers).
// RED_HIST: the three histogram cells are updated with fire-and-forget reductions instead of load-add-store (the thread does
// not wait for the loads; one thread per (d, t) and stream order keep every cell's additions in the reference's order).
template <int RFIX, bool RED_HIST = false, typename GetM, typename GetE, typename GetS, typename GetL>
__device__ __forceinline__ void fold_one(const ModelView &m, const StatsView &st, int64_t d, int t, int with_overlap,
                                         GetM get_m, GetE get_e, GetS get_ds, GetL get_dl, double *sums_local = nullptr) {
    const float nf = (float)m.N;
    // sums_local: the 11 running sums of (d, t) kept densely by the caller (a kernel that folds many sweeps of one (d, t) loads
    // them once and stores them once)
    double *sums = sums_local ? sums_local : st.sums + d * 11 * m.T + t;
    const int64_t ST = sums_local ? 1 : m.T;
    // simulation/mod.rs:555-578: replica-inner order
    double s0 = sums[0], s1 = sums[1 * ST], s2 = sums[2 * ST], s3 = sums[3 * ST], s4 = sums[4 * ST];
    const int R = RFIX > 0 ? RFIX : m.R, P = RFIX > 0 ? RFIX / 2 : m.P;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float mag = __fdiv_rn((float)get_m(r), nf);
        const float m2 = __fmul_rn(mag, mag);
        const float m4 = __fmul_rn(m2, m2);
        const float e = get_e(r);
        s0 = __dadd_rn(s0, (double)mag);
        s1 = __dadd_rn(s1, (double)m2);
        s2 = __dadd_rn(s2, (double)m4);
        s3 = __dadd_rn(s3, (double)e);
        s4 = __dadd_rn(s4, __dmul_rn((double)e, (double)e));  // stats.rs:23: powi(2) in f64
    }
    sums[0] = s0; sums[1 * ST] = s1; sums[2 * ST] = s2; sums[3 * ST] = s3; sums[4 * ST] = s4;
    if (!with_overlap) return;
    // statistics/overlap.rs:283-306 + :318-324, pair order
    const float nb = (float)(m.N * m.z);
    const int64_t bins = m.N + 1;
    double o0 = sums[5 * ST], o1 = sums[6 * ST], o2 = sums[7 * ST], o3 = sums[8 * ST], o4 = sums[9 * ST],
           o5 = sums[10 * ST];
#pragma unroll
    for (int p = 0; p < P; p++) {
        const long long dsp = get_ds(p);
        const long long dlk = get_dl(p);
        const float ql = __fdiv_rn((float)dlk, nb);
        const float q = __fdiv_rn((float)dsp, nf);
        const float q2 = __fmul_rn(q, q);
        const float ql2 = __fmul_rn(ql, ql);
        o0 = __dadd_rn(o0, (double)q);
        o1 = __dadd_rn(o1, (double)q2);
        o2 = __dadd_rn(o2, (double)__fmul_rn(q2, q2));
        o3 = __dadd_rn(o3, (double)ql);
        o4 = __dadd_rn(o4, (double)ql2);
        o5 = __dadd_rn(o5, (double)__fmul_rn(ql2, ql2));
        const int64_t idx = (dsp + m.N) / 2;
        const int64_t h = (d * m.T + t) * bins + idx;
        if (RED_HIST) {
            atomicAdd(st.hist + h, 1u);
            atomicAdd(st.ql_at_q + h, (double)ql);
            atomicAdd(st.ql2_at_q + h, (double)ql2);
        } else {
            st.hist[h] += 1u;
            st.ql_at_q[h] = __dadd_rn(st.ql_at_q[h], (double)ql);
            st.ql2_at_q[h] = __dadd_rn(st.ql2_at_q[h], (double)ql2);
        }
    }
    sums[5 * ST] = o0; sums[6 * ST] = o1; sums[7 * ST] = o2; sums[8 * ST] = o3; sums[9 * ST] = o4;
    sums[10 * ST] = o5;
}

// The same fold for a BATCH of up to FOLD_K recorded sweeps of one (d, t), by one warp.  Phase A: lane k derives the terms of
// sweep k (the f32 arithmetic of fold_one, widened to f64) into a shared-memory table — the divisions of the sweeps run side by side
// instead of one after the other.  Phase B: lane j < 11 adds column j to running sum j, sweeps in order, replicas / pairs in order,
// so every sum receives exactly fold_one's additions in fold_one's order; lane 11 walks the histogram cells in the same order.
// A single thread folding one sweep costs ~4000 cycles of dependent instructions; this costs a few hundred per sweep.
// Mv / Ev: [n][8] by replica, Sv / Lv: [n][4] by pair; tab: [FOLD_K][8][11] doubles; sums: the 11 dense running sums of (d, t).
constexpr int FOLD_K = 16;
__device__ __forceinline__ void fold_batch(const ModelView &m, const StatsView &st, int64_t d, int t, int lane, int n,
                                           const long long (*Mv)[8], const float (*Ev)[8], const long long (*Sv)[4],
                                           const long long (*Lv)[4], double (*tab)[8][11], double *sums) {
    const float nf = (float)m.N, nb = (float)(m.N * m.z);
    if (lane < n) {
        for (int r = 0; r < m.R; r++) {
            const float mag = __fdiv_rn((float)Mv[lane][r], nf);
            const float m2 = __fmul_rn(mag, mag);
            const float e = Ev[lane][r];
            double *o = tab[lane][r];
            o[0] = (double)mag;
            o[1] = (double)m2;
            o[2] = (double)__fmul_rn(m2, m2);
            o[3] = (double)e;
            o[4] = __dmul_rn((double)e, (double)e);  // stats.rs:23: powi(2) in f64
        }
        for (int p = 0; p < m.P; p++) {
            const float ql = __fdiv_rn((float)Lv[lane][p], nb);
            const float q = __fdiv_rn((float)Sv[lane][p], nf);
            const float q2 = __fmul_rn(q, q);
            const float ql2 = __fmul_rn(ql, ql);
            double *o = tab[lane][p];
            o[5] = (double)q;
            o[6] = (double)q2;
            o[7] = (double)__fmul_rn(q2, q2);
            o[8] = (double)ql;
            o[9] = (double)ql2;
            o[10] = (double)__fmul_rn(ql2, ql2);
        }
    }
    __syncwarp();
    if (lane < 11 && (lane < 5 || m.P > 0)) {
        const int cnt = lane < 5 ? m.R : m.P;
        double acc = sums[lane];
        for (int k = 0; k < n; k++)
            for (int i = 0; i < cnt; i++) acc = __dadd_rn(acc, tab[k][i][lane]);
        sums[lane] = acc;
    } else if (lane == 11 && m.P > 0) {
        const int64_t bins = m.N + 1;
        for (int k = 0; k < n; k++)
            for (int p = 0; p < m.P; p++) {
                const int64_t h = (d * m.T + t) * bins + (Sv[k][p] + m.N) / 2;
                atomicAdd(st.hist + h, 1u);
                atomicAdd(st.ql_at_q + h, tab[k][p][8]);
                atomicAdd(st.ql2_at_q + h, tab[k][p][9]);
            }
    }
    __syncwarp();
}

// The same fold with one THREAD per running sum: thread j < 5 owns sum j of simulation/mod.rs:555-578, thread 5 + k sum k of
// statistics/overlap.rs:283-306, thread 11 the histogram cells (12 consecutive threads per (d, t); the callers spread the (d, t)
// over the block).  One uniform loop: every thread derives the terms of replica i and pair i and adds its own, so each sum receives
// exactly fold_one's additions in fold_one's order — a thread's chain is R (or P) additions long instead of 5 R + 6 P.
template <bool SMALL = false, typename GetM, typename GetE, typename GetS, typename GetL>
__device__ __forceinline__ void fold_spread(const ModelView &m, const StatsView &st, int64_t d, int t, int j, GetM get_m, GetE get_e,
                                            GetS get_ds, GetL get_dl) {
    const float nf = (float)m.N, nb = (float)(m.N * m.z);
    const int64_t bins = m.N + 1;
    const bool owns = j < 5 || (j < 11 && m.P > 0);
    double *sum = st.sums + d * 11 * m.T + t + (int64_t)(j < 11 ? j : 0) * m.T;
    double acc = owns ? *sum : 0.0;
    const int n = m.R > m.P ? m.R : m.P;
    // SMALL: every integer input fits 32 bits (systems that live in shared memory): same value, a one-instruction conversion
    auto to_f = [](long long x) { return SMALL ? (float)(int)x : (float)x; };
    for (int i = 0; i < n; i++) {
        const bool hr = i < m.R, hp = i < m.P;
        const float mag = __fdiv_rn(to_f(get_m(hr ? i : 0)), nf);
        const float m2 = __fmul_rn(mag, mag);
        const float e = get_e(hr ? i : 0);
        const long long dsp = m.P > 0 ? get_ds(hp ? i : 0) : 0ll;
        const long long dlk = m.P > 0 ? get_dl(hp ? i : 0) : 0ll;
        const float ql = __fdiv_rn(to_f(dlk), nb);
        const float q = __fdiv_rn(to_f(dsp), nf);
        const float q2 = __fmul_rn(q, q);
        const float ql2 = __fmul_rn(ql, ql);
        float tf = mag;  // the thread's f32 term, widened once (every term but e^2 is an f32 value)
        tf = j == 1 ? m2 : tf;
        tf = j == 2 ? __fmul_rn(m2, m2) : tf;
        tf = j == 3 || j == 4 ? e : tf;
        tf = j == 5 ? q : tf;
        tf = j == 6 ? q2 : tf;
        tf = j == 7 ? __fmul_rn(q2, q2) : tf;
        tf = j == 8 ? ql : tf;
        tf = j == 9 ? ql2 : tf;
        tf = j == 10 ? __fmul_rn(ql2, ql2) : tf;
        const double td = (double)tf;
        const double term = j == 4 ? __dmul_rn(td, td) : td;  // stats.rs:23: powi(2) in f64
        if (owns && (j < 5 ? hr : hp)) acc = __dadd_rn(acc, term);
        if (j == 11 && hp) {  // the histogram cells in pair order (two pairs of a sweep may hit the same cell)
            const int64_t h = (d * m.T + t) * bins + (dsp + m.N) / 2;
            atomicAdd(st.hist + h, 1u);
            atomicAdd(st.ql_at_q + h, (double)ql);
            atomicAdd(st.ql2_at_q + h, (double)ql2);
        }
    }
    if (owns) *sum = acc;
}

// Same fold as fire-and-forget reductions (RED.ADD.F64 / RED.ADD.U32): no load, so the warp that finishes 32
// realizations in the fused msc3d epilogue never waits on memory.  Every accumulator is touched by exactly one thread
// per kernel and kernels of one chunk are stream-ordered, so each f64 cell still receives the reference's additions in
// the reference's order (IEEE round-to-nearest adds, same-address reductions of one thread stay in program order):
// the sums are bit-identical to the sequential fold.
template <int R, int NP>
__device__ __forceinline__ void fold_red(const ModelView &m, const StatsView &st, int64_t d, int t, const long long *Mv,
                                         const float *Ev, const long long *Sv, const long long *Lv) {
    const float nf = (float)m.N;
    double *sums = st.sums + d * 11 * m.T + t;
#pragma unroll
    for (int r = 0; r < R; r++) {  // simulation/mod.rs:555-578
        const float mag = __fdiv_rn((float)Mv[r], nf);
        const float m2 = __fmul_rn(mag, mag);
        const float m4 = __fmul_rn(m2, m2);
        const float e = Ev[r];
        atomicAdd(sums + 0 * m.T, (double)mag);
        atomicAdd(sums + 1 * m.T, (double)m2);
        atomicAdd(sums + 2 * m.T, (double)m4);
        atomicAdd(sums + 3 * m.T, (double)e);
        atomicAdd(sums + 4 * m.T, __dmul_rn((double)e, (double)e));
    }
    const float nb = (float)(m.N * m.z);
    const int64_t bins = m.N + 1;
#pragma unroll
    for (int p = 0; p < NP; p++) {  // statistics/overlap.rs:283-306 + :318-324
        const float ql = __fdiv_rn((float)Lv[p], nb);
        const float q = __fdiv_rn((float)Sv[p], nf);
        const float q2 = __fmul_rn(q, q);
        const float ql2 = __fmul_rn(ql, ql);
        atomicAdd(sums + 5 * m.T, (double)q);
        atomicAdd(sums + 6 * m.T, (double)q2);
        atomicAdd(sums + 7 * m.T, (double)__fmul_rn(q2, q2));
        atomicAdd(sums + 8 * m.T, (double)ql);
        atomicAdd(sums + 9 * m.T, (double)ql2);
        atomicAdd(sums + 10 * m.T, (double)__fmul_rn(ql2, ql2));
        const int64_t h = (d * m.T + t) * bins + (Sv[p] + m.N) / 2;
        atomicAdd(st.hist + h, 1u);
        atomicAdd(st.ql_at_q + h, (double)ql);
        atomicAdd(st.ql2_at_q + h, (double)ql2);
    }
}

// standalone: one thread per (realization d, temperature t), inputs from the per-system arrays
__global__ void fold_kernel(ModelView m, StatsView st, int with_overlap) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m.D * m.T) return;
    const int64_t d = gid / m.T;
    const int t = (int)(gid % m.T);
    fold_one<0>(
        m, st, d, t, with_overlap,
        [&](int r) { return m.mags[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
        [&](int r) { return m.energies[d * m.S + m.system_ids[d * m.S + r * m.T + t]]; },
        [&](int p) { return st.dot_spin[(d * m.P + p) * m.T + t]; },
        [&](int p) { return st.dot_link[(d * m.P + p) * m.T + t]; });
}

// ------------------------------------------------------------------------------------------------
// Integrated autocorrelation times of m^2 and q^2 (statistics/autocorrelation.rs, ring backend; driven from
// simulation/mod.rs:341-371, 551-594, 825-832).  One accumulator per (realization d, temperature t):
// ring f32 [K + 1], sum_prod f64 [K + 1], sum_o / sum_o2 f64.  All f64 operations are the reference's, in its order.
struct AutocorrView {
    int K;             // max lag (already clamped, mod.rs:342-344)
    float *ring;       // [D][T][K + 1]
    double *sum_prod;  // [D][T][K + 1]
    double *sum_o, *sum_o2;  // [D][T]
};

// After the fold of one recorded sweep (pre-exchange system_ids): block (d, t) rebuilds this sweep's two measurements from
// the per-system magnetisation sums and the pair dots -- m2 = (sum_r (f64)(f32)mag_r^2) / R (mod.rs:568-586),
// q2 = (sum_p (f64)(f32)q_p^2) / P (overlap.rs:314-316, mod.rs:588-594) -- and pushes them (autocorrelation.rs:68-112).
__global__ void __launch_bounds__(64)
autocorr_push_kernel(ModelView m, const long long *dot_spin, AutocorrView am, AutocorrView aq, long long n_recorded) {
    __shared__ float o_sh[2];
    const int64_t dt = blockIdx.x;
    const int64_t d = dt / m.T;
    const int t = (int)(dt % m.T);
    const float nf = (float)m.N;
    if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int r = 0; r < m.R; r++) {
            const float mag = __fdiv_rn((float)m.mags[d * m.S + m.system_ids[d * m.S + r * m.T + t]], nf);
            acc = __dadd_rn(acc, (double)__fmul_rn(mag, mag));
        }
        o_sh[0] = (float)__dmul_rn(acc, __ddiv_rn(1.0, (double)m.R));
    } else if (threadIdx.x == 32 && aq.ring) {
        double acc = 0.0;
        for (int p = 0; p < m.P; p++) {
            const float q = __fdiv_rn((float)dot_spin[(d * m.P + p) * m.T + t], nf);
            acc = __dadd_rn(acc, (double)__fmul_rn(q, q));
        }
        o_sh[1] = (float)__dmul_rn(acc, __ddiv_rn(1.0, (double)m.P));
    }
    __syncthreads();
    for (int which = 0; which < 2; which++) {
        const AutocorrView &a = which == 0 ? am : aq;
        if (!a.ring) continue;
        const int len = a.K + 1, pos = (int)(n_recorded % len);
        const long long n_back = n_recorded < a.K ? n_recorded : a.K;
        const double o = (double)o_sh[which];
        float *ring = a.ring + dt * len;
        double *sp = a.sum_prod + dt * len;
        for (long long delta = threadIdx.x; delta <= n_back; delta += blockDim.x) {
            const float other = delta == 0 ? o_sh[which] : ring[pos >= delta ? pos - delta : pos + len - delta];
            sp[delta] = __dadd_rn(sp[delta], __dmul_rn(o, (double)other));
        }
        if (threadIdx.x == 0) {
            a.sum_o[dt] = __dadd_rn(a.sum_o[dt], o);
            a.sum_o2[dt] = __dadd_rn(a.sum_o2[dt], __dmul_rn(o, o));
        }
    }
    __syncthreads();  // every lag has read the old ring before the newest value lands (the slot at pos is never read above)
    if (threadIdx.x == 0) am.ring[dt * (am.K + 1) + (int)(n_recorded % (am.K + 1))] = o_sh[0];
    if (threadIdx.x == 32 && aq.ring) aq.ring[dt * (aq.K + 1) + (int)(n_recorded % (aq.K + 1))] = o_sh[1];
}

// finish() + sokal_tau (autocorrelation.rs:114-124, 166-210): one thread per (d, t)
__global__ void autocorr_tau_kernel(int64_t n_dt, AutocorrView a, long long n_recorded, double *tau_out) {
    const int64_t dt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dt >= n_dt) return;
    const int len = a.K + 1;
    bool degenerate = n_recorded == 0;
    double mm = 0.0, var = 0.0;
    if (!degenerate) {
        const double mcount = (double)n_recorded;
        const double mean = __ddiv_rn(a.sum_o[dt], mcount);
        mm = __dmul_rn(mean, mean);
        var = __dsub_rn(__ddiv_rn(a.sum_o2[dt], mcount), mm);
        if (var <= 0.0) degenerate = true;
    }
    double tau = 0.5;
    for (int w = 1; w < len; w++) {
        double g = 0.0;
        const long long cnt = n_recorded > w ? n_recorded - w : 0;
        if (!degenerate && cnt > 0) g = __ddiv_rn(__dsub_rn(__ddiv_rn(a.sum_prod[dt * len + w], (double)cnt), mm), var);
        tau = __dadd_rn(tau, g);
        if ((double)w >= __dmul_rn(5.0, tau)) break;
    }
    tau_out[dt] = tau;
}

// Equilibration diagnostic (statistics/equilibration.rs; simulation/mod.rs:511-541): after EVERY sweep, thread (d, t) adds the
// replica-mean energy and the pair-mean link overlap at slot t (f32 means, f64 running sums) and, at a checkpoint, stores
// the running averages: snap[(d * n_ckpt + ckpt) * 2 * T + {0, 1} * T + t].
__global__ void equil_push_kernel(ModelView m, const long long *dot_link, double *sum_e, double *sum_ql, long long count_after,
                                  int ckpt, int n_ckpt, double *snap) {
    const int64_t dt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dt >= m.D * m.T) return;
    const int64_t d = dt / m.T;
    const int t = (int)(dt % m.T);
    float e = 0.0f;
    for (int r = 0; r < m.R; r++) e = __fadd_rn(e, m.energies[d * m.S + m.system_ids[d * m.S + r * m.T + t]]);
    e = __fmul_rn(e, __fdiv_rn(1.0f, (float)m.R));
    float ql = 0.0f;
    if (m.P > 0) {
        const float nb = (float)(m.N * m.z);
        for (int p = 0; p < m.P; p++) ql = __fadd_rn(ql, __fdiv_rn((float)dot_link[(d * m.P + p) * m.T + t], nb));
        ql = __fmul_rn(ql, __fdiv_rn(1.0f, (float)m.P));
    }
    const double se = __dadd_rn(sum_e[dt], (double)e), sq = __dadd_rn(sum_ql[dt], (double)ql);
    sum_e[dt] = se;
    sum_ql[dt] = sq;
    if (ckpt >= 0) {
        const double c = (double)count_after;
        double *o = snap + ((size_t)d * n_ckpt + ckpt) * 2 * m.T;
        o[t] = __ddiv_rn(se, c);
        o[m.T + t] = __ddiv_rn(sq, c);
    }
}

// realization.rs:109-120
__device__ __forceinline__ void pt_record_arrival(const PtView &pt, int64_t base, int system, int slot) {
    if (slot == pt.hot_slot) {
        if (pt.trip_state[base + system] == 2) pt.round_trips[base + system] += 1ull;
        pt.trip_state[base + system] = 1;
        return;
    }
    if (slot == pt.cold_slot && pt.trip_state[base + system] == 1) pt.trip_state[base + system] = 2;
}

// tempering.rs:73-102; logtab[draw] = host-libm logf(draw / 2^24) so the decision is the CPU's bit for bit
__device__ __forceinline__ bool pt_attempt_edge(const ModelView &m, const PtView &pt, int64_t d, int r, int edge,
                                                uint32_t draw) {
    int32_t *sid = m.system_ids + d * m.S + r * m.T;
    const float temp_1 = m.temps[edge], temp_2 = m.temps[edge + 1];
    const int left = sid[edge], right = sid[edge + 1];
    const float energy_1 = m.energies[d * m.S + left], energy_2 = m.energies[d * m.S + right];
    const float delta = __fmul_rn(__fmul_rn((float)m.N, __fsub_rn(energy_2, energy_1)),
                                  __fsub_rn(__fdiv_rn(1.0f, temp_1), __fdiv_rn(1.0f, temp_2)));
    const bool accepted = delta >= m.logtab[draw];
    const int64_t eb = d * (m.T - 1);
    // the T-1 edge counters of a realization are shared by its R replica threads
    atomicAdd(&pt.edge_attempts[eb + edge], 1ull);
    if (!accepted) return false;
    sid[edge] = right;
    sid[edge + 1] = left;
    atomicAdd(&pt.edge_acceptances[eb + edge], 1ull);
    pt_record_arrival(pt, d * m.S, left, edge + 1);   // realization.rs:80-81
    pt_record_arrival(pt, d * m.S, right, edge);
    return true;
}

// the same attempt with everything that does not depend on the energies prepared by the caller: log_u = logtab[draw] (fetched
// while the sweep ran), dbeta = 1 / T_edge - 1 / T_edge+1 (the two f32 divisions and the subtraction of the expression above)
__device__ __forceinline__ bool pt_attempt_edge_pre(const ModelView &m, const PtView &pt, int64_t d, int r, int edge, float log_u,
                                                    float dbeta) {
    int32_t *sid = m.system_ids + d * m.S + r * m.T;
    const int left = sid[edge], right = sid[edge + 1];
    const float energy_1 = m.energies[d * m.S + left], energy_2 = m.energies[d * m.S + right];
    const float delta = __fmul_rn(__fmul_rn((float)m.N, __fsub_rn(energy_2, energy_1)), dbeta);
    const bool accepted = delta >= log_u;
    const int64_t eb = d * (m.T - 1);
    atomicAdd(&pt.edge_attempts[eb + edge], 1ull);
    if (!accepted) return false;
    sid[edge] = right;
    sid[edge + 1] = left;
    atomicAdd(&pt.edge_acceptances[eb + edge], 1ull);
    pt_record_arrival(pt, d * m.S, left, edge + 1);   // realization.rs:80-81
    pt_record_arrival(pt, d * m.S, right, edge);
    return true;
}

// K8: one thread per (realization, replica).  RNG-SPEC PT domain: key = realization key,
// counter = {edge | 0xFFFFFFFF, pt_event, replica, TAG_PT}.
__device__ __forceinline__ void pt_exchange_body(const ModelView &m, const PtView &pt, const int64_t d, const int r, const int schedule,
                                                 const int first_parity, const uint32_t pt_event) {
    const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    if (schedule == 0) {  // tempering.rs:20-42
        const u32x4 o = philox4x32(0xFFFFFFFFu, pt_event, (uint32_t)r, TAG_PT, k0, k1);
        const int edge = (int)(((uint64_t)o.y * (uint64_t)(m.T - 1)) >> 32);
        pt_attempt_edge(m, pt, d, r, edge, o.x >> 8);
    } else {  // tempering.rs:45-70
        for (int pi = 0; pi < 2; pi++) {
            const int parity = pi == 0 ? first_parity : 1 - first_parity;
            for (int edge = parity; edge < m.T - 1; edge += 2) {
                const u32x4 o = philox4x32((uint32_t)edge, pt_event, (uint32_t)r, TAG_PT, k0, k1);
                pt_attempt_edge(m, pt, d, r, edge, o.x >> 8);
            }
        }
    }
}

__global__ void pt_exchange_kernel(ModelView m, PtView pt, int schedule, int first_parity, uint32_t pt_event) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m.D * m.R || m.T < 2) return;
    pt_exchange_body(m, pt, gid / m.R, (int)(gid % m.R), schedule, first_parity, pt_event);
}

// K8, multispin layout: one warp per (word group g, replica r) ladder, lane l = realization 32g + l.  Same decisions
// and counters as pt_exchange_kernel; the per-edge lane masks are produced with warp votes (plain stores, no
// pre-zeroing): swap_mask[(g*R + r)*(T-1) + edge] = lanes whose configurations cross that edge.
__global__ void __launch_bounds__(128) pt_exchange_msc_kernel(ModelView m, PtView pt, int schedule, int first_parity, uint32_t pt_event) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t G = (m.D + 31) >> 5;
    if (wid >= G * m.R || m.T < 2) return;
    const int64_t g = wid / m.R;
    const int r = (int)(wid % m.R);
    const int64_t d = g * 32 + lane;
    const bool live = d < m.D;
    uint32_t *mask = pt.swap_mask + wid * (m.T - 1);
    uint32_t k0 = 0, k1 = 0;
    if (live) {
        const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
        k0 = (uint32_t)key;
        k1 = (uint32_t)(key >> 32);
    }
    if (schedule == 0) {  // tempering.rs:20-42: every lane draws its own edge
        for (int e = lane; e < m.T - 1; e += 32) mask[e] = 0u;
        __syncwarp();
        int edge = -1;
        bool acc = false;
        if (live) {
            const u32x4 o = philox4x32(0xFFFFFFFFu, pt_event, (uint32_t)r, TAG_PT, k0, k1);
            edge = (int)(((uint64_t)o.y * (uint64_t)(m.T - 1)) >> 32);
            acc = pt_attempt_edge(m, pt, d, r, edge, o.x >> 8);
        }
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, edge);
        const uint32_t accepted = __ballot_sync(0xFFFFFFFFu, acc);
        if (live && lane == __ffs(same) - 1) mask[edge] = same & accepted;
    } else {  // tempering.rs:45-70
        for (int pi = 0; pi < 2; pi++) {
            const int parity = pi == 0 ? first_parity : 1 - first_parity;
            for (int edge = parity; edge < m.T - 1; edge += 2) {
                bool acc = false;
                if (live) {
                    const u32x4 o = philox4x32((uint32_t)edge, pt_event, (uint32_t)r, TAG_PT, k0, k1);
                    acc = pt_attempt_edge(m, pt, d, r, edge, o.x >> 8);
                }
                const uint32_t accepted = __ballot_sync(0xFFFFFFFFu, acc);
                if (lane == 0) mask[edge] = accepted;
            }
        }
    }
}

// overlap.rs:128-142: per-(t, bin) sums over realizations, in realization order
__global__ void reduce_hist_kernel(ModelView m, StatsView st, unsigned long long *hist_out, double *ql_out, double *ql2_out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)m.T * (m.N + 1);
    if (gid >= per) return;
    unsigned long long h = 0;
    double a = 0.0, b = 0.0;
    for (int64_t d = 0; d < m.D; d++) {
        h += st.hist[d * per + gid];
        a = __dadd_rn(a, st.ql_at_q[d * per + gid]);
        b = __dadd_rn(b, st.ql2_at_q[d * per + gid]);
    }
    hist_out[gid] = h;
    ql_out[gid] = a;
    ql2_out[gid] = b;
}

__global__ void widen_u32_kernel(const uint32_t *in, unsigned long long *out, int64_t n) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) out[gid] = in[gid];
}

}  // namespace pp
