// pp_kernels_slab.cuh — one large 3-D hypercubic ferromagnet, stride geometry, slab-decomposed along x0.
//
// The table-driven int8 kernels need 8*z' bytes of neighbour indices per site (24 GiB at 1024^3, like the
// reference's Lattice: spin-sim/src/geometry/lattice.rs:63-82).  This path has no tables: a thread owns the 8
// consecutive sites of one row segment, its neighbours are the same segment in the rows x1 +- 1 and the planes
// x0 +- 1 plus one edge byte, and all six neighbour counts of the 8 sites come from five 64-bit adds (SWAR).
//
// Replaces, for this layout,
//   metropolis_sweep / gibbs_sweep (lookup rule)     spin-sim/src/mcmc/sweep.rs:170-185, 220-284
//   compute_energies_and_magnetizations_into         spin-sim/src/spins/energy.rs:59-110
//   Realization::new / reset (spin draw)             spin-sim/src/simulation/realization.rs:177-182
//
// Storage: u8 [S][P + 2][L1][L2], one byte per spin, 1 = spin -1, 0 = spin +1.  Planes 1..P are the slab's own
// planes (global x0 = first_plane + p - 1), planes 0 and P + 1 are halos holding the neighbouring slabs' boundary
// planes (the periodic images when there is one slab).  P is even, so the checkerboard colour (x0 + x1 + x2) & 1 of
// a site does not depend on how the lattice is cut.  Draws follow RNG-SPEC exactly as the table-driven int8 kernels
// (pp_kernels_int8.cuh): colour rank = site >> 1, one Philox call per 8-site segment and colour, stream = system id.
#pragma once
#include "pp_device.cuh"

namespace pp {

struct SlabView {
    uint8_t *spins;            // [S][P + 2][L1][L2]
    int P, L1, L2;
    int64_t plane;             // L1 * L2
    int64_t sys_stride;        // (P + 2) * plane
    int64_t first_plane;       // global x0 of local plane 1
    int64_t chunks_per_plane;  // L1 * L2 / 8
    int kpr_shift;             // log2(L2 / 8) when that is a power of two, else -1
    int tile_x, tile_x_shift;  // sweep tile: tile_x (a power of two <= 32) segments of a row x 256 / tile_x rows
    uint32_t k0, k1;           // Philox key of the realization (realization_seed(seed, sample_offset), set per launch)
};

// segment index inside a plane -> (row x1, segment k of the row)
__device__ __forceinline__ void slab_split(const SlabView &v, const int64_t c, int &x1, int &k) {
    const int kpr = v.L2 >> 3;
    if (v.kpr_shift >= 0) {
        x1 = (int)(c >> v.kpr_shift);
        k = (int)c & (kpr - 1);
    } else {
        x1 = (int)((uint32_t)c / (uint32_t)kpr);
        k = (int)c - x1 * kpr;
    }
}

__device__ __forceinline__ uint64_t ld8(const uint8_t *p) {
    const uint2 v = *reinterpret_cast<const uint2 *>(p);
    return (uint64_t)v.x | ((uint64_t)v.y << 32);
}

constexpr int SLAB_PR = 8;  // planes a sweep thread marches through

// One colour half-step of the local planes [pa, pa + np) (and, when pb > 0, of plane pb as one more range: the two
// boundary planes go in one launch).  grid = (column tiles, ceil(np / SLAB_PR) [+ 1], slots); block = 256
// threads = a tile of tile_x row segments x 256 / tile_x rows (tile_x = v.tile_x).  A thread owns one column of segments
// (x1, k) and marches through SLAB_PR planes with a three-plane register window, so the x0 neighbours cost one new
// 8-byte load per update and the x1 neighbours are segments other threads of the same tile load at the same time (L1).
__global__ void __launch_bounds__(256)
slab_sweep_kernel(ModelView m, SlabView v, int colour, uint32_t sweep_index, int pa, int np, int pb) {
    __shared__ uint32_t thr[8];
    const int slot = blockIdx.z;
    const int t = slot % m.T;  // realization.rs:166: temperatures repeat with period T
    if (threadIdx.x < 7) thr[threadIdx.x] = m.lut[t * 13 + 2 * threadIdx.x];  // sweep.rs:162-166, index ec + 2z' = 2 * unsat
    __syncthreads();
    const int kpr = v.L2 >> 3;
    const int tiles_x = (kpr + v.tile_x - 1) / v.tile_x, tile_y = 256 / v.tile_x;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int k = tx * v.tile_x + (threadIdx.x & (v.tile_x - 1));
    const int x1 = ty * tile_y + (threadIdx.x >> v.tile_x_shift);
    if (k >= kpr || x1 >= v.L1) return;
    const int n_ranges = (np + SLAB_PR - 1) / SLAB_PR;
    const bool second = (int)blockIdx.y >= n_ranges;  // the extra range: plane pb alone
    const int p0 = second ? pb : pa + blockIdx.y * SLAB_PR, p1 = second ? pb + 1 : min(p0 + SLAB_PR, pa + np);
    const uint32_t sys = (uint32_t)m.system_ids[slot];  // parallel.rs:27-33: spins by system, temperature by slot
    uint8_t *col = v.spins + (int64_t)sys * v.sys_stride + (int64_t)x1 * v.L2 + 8 * k;  // plane 0 of this column
    const int x1m = x1 ? x1 - 1 : v.L1 - 1, x1p = x1 + 1 == v.L1 ? 0 : x1 + 1;
    const int64_t dym = (int64_t)(x1m - x1) * v.L2, dyp = (int64_t)(x1p - x1) * v.L2;
    const int xe_l = k ? -1 : v.L2 - 1, xe_r = k + 1 == kpr ? 8 - v.L2 : 8;  // byte left / right of the segment
    uint64_t Xm = ld8(col + (int64_t)(p0 - 1) * v.plane), C = ld8(col + (int64_t)p0 * v.plane);
    for (int p = p0; p < p1; p++) {
        uint8_t *row = col + (int64_t)p * v.plane;
        const uint64_t Xp = ld8(row + v.plane);
        const uint64_t Ym = ld8(row + dym), Yp = ld8(row + dyp);
        const uint32_t gx0 = (uint32_t)(v.first_plane + p - 1);
        const uint32_t off = (colour ^ gx0 ^ x1) & 1u;  // active sites of the segment: x2 = 8k + 2l + off
        // off = 0: site 0 needs the byte left of the segment; off = 1: site 7 needs the byte right of it
        const uint64_t E = row[off ? xe_r : xe_l];
        const uint64_t left = (C << 8) | (off ? 0ull : E), right = (C >> 8) | (off ? E << 56 : 0ull);
        const uint64_t down = Xm + Xp + Ym + Yp + left + right;  // per byte: down-spin neighbours (<= 6, no carries)
        const uint32_t q = (gx0 * (uint32_t)v.L1 + (uint32_t)x1) * (uint32_t)kpr + (uint32_t)k;  // segment index = colour rank >> 2
        const u32x4 o = philox4x32(q, sweep_index, sys, TAG_SWEEP | (uint32_t)colour, v.k0, v.k1);
        // ferromagnet: a bond is unsatisfied iff the two spins differ, so per byte unsat = s ? 6 - down : down
        // = (down ^ 7s) - s for all 8 sites at once; then the four active sites are brought to the even bytes and the two
        // 32-bit halves are read with constant shifts
        const uint64_t unsat8 = (down ^ (C * 7ull)) - C;
        const uint64_t ush = unsat8 >> (8 * off);
        const uint32_t u2[2] = {(uint32_t)ush, (uint32_t)(ush >> 32)};
        uint32_t f2[2] = {0u, 0u};
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const uint32_t unsat = (u2[l >> 1] >> (16 * (l & 1))) & 0xFFu;
            if ((pick(o, l) >> 8) < thr[unsat]) f2[l >> 1] |= 1u << (16 * (l & 1));  // sweep.rs:182-184
        }
        const uint64_t out = C ^ (((uint64_t)f2[0] | ((uint64_t)f2[1] << 32)) << (8 * off));
        *reinterpret_cast<uint2 *>(row) = make_uint2((uint32_t)out, (uint32_t)(out >> 32));
        Xm = C;  // the neighbours' bytes of the other colour are what the next plane reads: unchanged by this update
        C = Xp;
    }
}

// K0: spin -1 iff the INIT-domain draw < 2^23 (realization.rs:180); grid = (chunk blocks, P, S)
__global__ void __launch_bounds__(256) slab_init_kernel(ModelView m, SlabView v) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.chunks_per_plane) return;
    const int p = blockIdx.y + 1;
    const uint32_t sys = blockIdx.z;
    const uint64_t seg = ((uint64_t)(v.first_plane + p - 1)) * v.chunks_per_plane + c;  // global site index >> 3
    uint64_t out = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const u32x4 o = philox4x32((uint32_t)(2 * seg + h), 0u, sys, TAG_INIT, v.k0, v.k1);
#pragma unroll
        for (int l = 0; l < 4; l++)
            if ((pick(o, l) >> 8) < (1u << 23)) out |= 1ull << (8 * (4 * h + l));
    }
    uint8_t *dst = v.spins + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane + 8 * c;
    *reinterpret_cast<uint2 *>(dst) = make_uint2((uint32_t)out, (uint32_t)(out >> 32));
}

// K5: per-system partial sums of this slab: unsatisfied forward bonds and down spins (energy.rs:92-109 counts every
// bond once through its forward direction).  grid = (chunk blocks, P, S); partial[2*sys] += unsat, [2*sys+1] += down.
__global__ void __launch_bounds__(256) slab_energy_kernel(SlabView v, unsigned long long *partial) {
    __shared__ long long sh[32];
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y + 1;
    const uint32_t sys = blockIdx.z;
    long long unsat = 0, dn = 0;
    if (c < v.chunks_per_plane) {
        int x1, k;
        slab_split(v, c, x1, k);
        const uint8_t *row = v.spins + (int64_t)sys * v.sys_stride + (int64_t)p * v.plane + (int64_t)x1 * v.L2;
        const int x1p = x1 + 1 == v.L1 ? 0 : x1 + 1;
        const uint64_t C = ld8(row + 8 * k), Xp = ld8(row + v.plane + 8 * k), Yp = ld8(row + (int64_t)(x1p - x1) * v.L2 + 8 * k);
        const uint64_t E = row[8 * k + 8 == v.L2 ? 0 : 8 * k + 8];
        const uint64_t right = (C >> 8) | (E << 56);
        unsat = __popcll(C ^ Xp) + __popcll(C ^ Yp) + __popcll(C ^ right);
        dn = __popcll(C);
    }
    const long long tu = block_sum<long long>(unsat, sh);
    const long long td = block_sum<long long>(dn, sh);
    if (threadIdx.x == 0) {
        atomicAdd(&partial[2 * sys], (unsigned long long)tu);
        atomicAdd(&partial[2 * sys + 1], (unsigned long long)td);
    }
}

// totals (summed over slabs) -> e = (sum_bonds s s) / N with the reference's final f32 division (energy.rs:108) and M
__global__ void slab_finish_energy_kernel(ModelView m, const unsigned long long *total, int want_mags) {
    const int sys = blockIdx.x * blockDim.x + threadIdx.x;
    if (sys >= m.S) return;
    const long long bonds = 3ll * m.N - 2ll * (long long)total[2 * sys];
    m.energies[sys] = __fdiv_rn((float)bonds, (float)m.N);
    if (want_mags) m.mags[sys] = (long long)m.N - 2ll * (long long)total[2 * sys + 1];
}

// own planes <-> +-1 int8 in the reference's order [S][planes][L1][L2]; dir 0: unpack to ext, 1: pack from ext.
// ext_sys_stride / ext_off place this slab's planes inside the caller's array (whole lattice or local planes only).
__global__ void slab_convert_kernel(SlabView v, int8_t *ext, int64_t ext_sys_stride, int64_t ext_off, int dir) {
    const int64_t per_sys = (int64_t)v.P * v.plane;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sys = blockIdx.y;
    if (i >= per_sys) return;
    uint8_t *b = v.spins + (int64_t)sys * v.sys_stride + v.plane + i;
    int8_t *e = ext + (int64_t)sys * ext_sys_stride + ext_off + i;
    if (dir == 0) *e = *b ? (int8_t)-1 : (int8_t)1;
    else *b = *e < 0 ? 1 : 0;
}

}  // namespace pp
