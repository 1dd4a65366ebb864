// pp_kernels_int8.cuh — int8-per-spin layout, any lattice / coupling class / sweep mode.
//
// K0 init_spins          <- spin-sim/src/simulation/realization.rs:177-182
// K1/K3/K4 sweep_colour  <- spin-sim/src/mcmc/sweep.rs:8-19 (local field), :170-185 (lookup rule),
//                           :35-48 + :256 (Metropolis log form), :279-282 (Gibbs log form)
// K5 energy_mag          <- spin-sim/src/spins/energy.rs:78-110
// K6 overlap_dots        <- spin-sim/src/statistics/overlap.rs:259-281
//
// One launch per colour class: same-colour sites never interact, so the order inside a launch is
// irrelevant and the launch sequence (colour 0, 1, ...) is the RNG-SPEC visit order.
#pragma once
#include "pp_device.cuh"

namespace pp {

constexpr int SWEEP_BLOCK = 128;

// ---------------------------------------------------------------------------------------------
// K0: +-1 from the INIT-domain draw of (realization key, system, site): draw24 < 2^23 -> -1
// grid.x = D*S systems, grid.y = quads of sites
__global__ void init_spins_int8_kernel(ModelView m) {
    int64_t sysg = blockIdx.x;  // d*S + sys
    int64_t d = sysg / m.S;
    uint32_t sys = (uint32_t)(sysg % m.S);
    int64_t q = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (q * 4 >= m.N) return;
    uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
    u32x4 o = philox4x32((uint32_t)q, 0u, sys, TAG_INIT, (uint32_t)key, (uint32_t)(key >> 32));
    int8_t *s = m.spins + sysg * m.N;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        int64_t i = q * 4 + l;
        if (i < m.N) s[i] = (pick(o, l) >> 8) < (1u << 23) ? (int8_t)-1 : (int8_t)1;
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K3/K4: one colour class of one sweep.  grid.x = D*S slots, grid.y = quads of colour ranks.
template <int CLASS>
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_colour_int8_kernel(ModelView m, int colour, uint32_t sweep_index, int sweep_mode, int exact_log) {
    const int64_t slotg = blockIdx.x;  // d*S + slot
    const int64_t d = slotg / m.S;
    const int slot = (int)(slotg % m.S);
    const uint32_t sys = (uint32_t)m.system_ids[slotg];  // parallel.rs:27-33: spins by system, temperature by slot
    const int t = slot % m.T;                            // realization.rs:166: temperatures repeat with period T
    const uint32_t cs = m.colour_start[colour], ce = m.colour_start[colour + 1];
    const uint32_t q = blockIdx.y * blockDim.x + threadIdx.x;
    if (cs + q * 4 >= ce) return;
    const uint64_t key = realization_seed(m.seed, (uint64_t)(m.sample_offset + d));
    const u32x4 o = philox4x32(q, sweep_index, sys, TAG_SWEEP | (uint32_t)colour, (uint32_t)key, (uint32_t)(key >> 32));
    int8_t *s = m.spins + (d * m.S + sys) * m.N;
    const int z = m.z, z2 = 2 * m.z;
    const int width = 4 * z + 1;
    const float temp = m.temps[t];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        uint32_t p = cs + q * 4 + l;
        if (p >= ce) break;
        const uint32_t i = m.order[p];
        const uint32_t draw = pick(o, l) >> 8;
        const uint32_t *nb = m.nbr + (size_t)i * z2;
        const int si = s[i];
        if (CLASS != COUP_F32) {
            int h = 0;
            if (CLASS == COUP_FERRO) {
                for (int k = 0; k < z2; k++) h += s[nb[k]];
            } else {
                const int8_t *J = m.J8 + (size_t)d * m.N * z;
                for (int k = 0; k < z; k++) {
                    uint32_t jf = nb[2 * k], jb = nb[2 * k + 1];
                    h += s[jf] * J[(size_t)i * z + k];
                    h += s[jb] * J[(size_t)jb * z + k];
                }
            }
            const int ec = -si * h;  // sweep.rs:178
            if (draw < m.lut[t * width + ec + 2 * z]) s[i] = (int8_t)-si;
        } else {
            const float *J = m.Jf + (size_t)d * m.N * z;
            float h = 0.0f;  // sweep.rs:10-17: d-major, forward then backward, no fused multiply-add
            for (int k = 0; k < z; k++) {
                uint32_t jf = nb[2 * k], jb = nb[2 * k + 1];
                h = __fadd_rn(h, __fmul_rn((float)s[jf], J[(size_t)i * z + k]));
                h = __fadd_rn(h, __fmul_rn((float)s[jb], J[(size_t)jb * z + k]));
            }
            const float eng_change = __fmul_rn(-(float)si, h);  // sweep.rs:43-44
            const float u = (float)draw * (1.0f / 16777216.0f);
            float lg;
            if (sweep_mode == 0)
                lg = exact_log ? m.logtab[draw] : logf(u);  // sweep.rs:256
            else
                lg = exact_log ? m.glogtab[draw] : logf(__fdiv_rn(u, __fsub_rn(1.0f, u)));  // sweep.rs:279-282
            const float thr = __fmul_rn(__fdiv_rn(temp, 2.0f), lg);
            if (eng_change >= thr) s[i] = (int8_t)-si;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5: energies (+ magnetisation sums), one block per (realization, system).
// e = (sum_i sum_d s_i s_fwd J) / N with the reference's final f32 division (energy.rs:108);
// unit couplings accumulate in int64 (exact wherever the reference's f32 sum is), fp32 couplings in f64.
template <int CLASS>
__global__ void __launch_bounds__(256) energy_mag_int8_kernel(ModelView m, int want_mags) {
    __shared__ long long sh_ll[32];
    __shared__ double sh_d[32];
    const int64_t sysg = blockIdx.x;  // d*S + sys
    const int64_t d = sysg / m.S;
    const int8_t *s = m.spins + sysg * m.N;
    const int z = m.z, z2 = 2 * m.z;
    long long acc_i = 0, acc_m = 0;
    double acc_f = 0.0;
    for (int64_t i = threadIdx.x; i < m.N; i += blockDim.x) {
        const int si = s[i];
        acc_m += si;
        const uint32_t *nb = m.nbr + (size_t)i * z2;
        if (CLASS == COUP_FERRO) {
            int h = 0;
            for (int k = 0; k < z; k++) h += s[nb[2 * k]];
            acc_i += si * h;
        } else if (CLASS == COUP_UNIT) {
            const int8_t *J = m.J8 + ((size_t)d * m.N + i) * z;
            int h = 0;
            for (int k = 0; k < z; k++) h += s[nb[2 * k]] * J[k];
            acc_i += si * h;
        } else {
            const float *J = m.Jf + ((size_t)d * m.N + i) * z;
            for (int k = 0; k < z; k++) acc_f += (double)((float)(si * s[nb[2 * k]]) * J[k]);
        }
    }
    if (CLASS == COUP_F32) {
        double tot = block_sum<double>(acc_f, sh_d);
        if (threadIdx.x == 0) m.energies[sysg] = __fdiv_rn((float)tot, (float)m.N);
    } else {
        long long tot = block_sum<long long>(acc_i, sh_ll);
        if (threadIdx.x == 0) m.energies[sysg] = __fdiv_rn((float)tot, (float)m.N);
    }
    if (want_mags) {
        long long tm = block_sum<long long>(acc_m, sh_ll);
        if (threadIdx.x == 0) m.mags[sysg] = tm;
    }
}

// ---------------------------------------------------------------------------------------------
// K6: integer overlap dots, one block per (realization, pair, temperature slot)
__global__ void __launch_bounds__(256) overlap_dots_int8_kernel(ModelView m, long long *dot_spin, long long *dot_link) {
    __shared__ long long sh[32];
    const int64_t idx = blockIdx.x;  // (d*P + p)*T + t
    const int t = (int)(idx % m.T);
    const int p = (int)((idx / m.T) % m.P);
    const int64_t d = idx / ((int64_t)m.T * m.P);
    const int sa = m.system_ids[d * m.S + (2 * p) * m.T + t];      // overlap.rs:263-264
    const int sb = m.system_ids[d * m.S + (2 * p + 1) * m.T + t];
    const int8_t *a = m.spins + (d * m.S + sa) * m.N;
    const int8_t *b = m.spins + (d * m.S + sb) * m.N;
    const int z = m.z, z2 = 2 * m.z;
    long long ds = 0, dl = 0;
    for (int64_t j = threadIdx.x; j < m.N; j += blockDim.x) {
        const int q = a[j] * b[j];
        ds += q;
        const uint32_t *nb = m.nbr + (size_t)j * z2;
        int acc = 0;
        for (int k = 0; k < z; k++) {
            uint32_t n = nb[2 * k];
            acc += a[n] * b[n];
        }
        dl += q * acc;
    }
    long long tds = block_sum<long long>(ds, sh);
    long long tdl = block_sum<long long>(dl, sh);
    if (threadIdx.x == 0) {
        dot_spin[idx] = tds;
        dot_link[idx] = tdl;
    }
}

}  // namespace pp
