// pp_plan.h — host-side lattice plan: neighbour offsets lowered to index/stride tables plus the
// colouring that fixes the checkerboard visit order.
//
// Geometry follows spin-sim/src/geometry/lattice.rs:44-93 (row-major sites, strides[d] =
// prod shape[d+1..], forward neighbour = +offset, backward = -offset, periodic wrap by rem_euclid).
// The colouring has no reference equivalent (the reference sweeps in typewriter order,
// spin-sim/src/mcmc/sweep.rs:51-97); its rule is part of RNG-SPEC:
//   1. the first linear colouring c(x) = (sum_d a_d x_d) mod m, m = 2..8, a in lexicographic order,
//      with m | a_d*shape_d for every d (well defined on the torus) and sum_d a_d o_d != 0 mod m
//      for every offset o;
//   2. otherwise greedy first-fit in site order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pp {

struct LatticePlan {
    int n_dims = 0, z = 0;  // z = number of forward directions (coordination 2z)
    int64_t n_spins = 0;
    std::vector<int64_t> shape, strides;
    std::vector<int64_t> offsets;  // [z][n_dims]
    bool hypercubic = false;       // offsets are the unit vectors in order
    // nbr[i*2z + 2d] = forward neighbour in direction d, nbr[i*2z + 2d + 1] = backward neighbour
    std::vector<uint32_t> nbr;
    // colouring
    int n_colours = 0;
    bool linear_colouring = false;
    int colour_mod = 0;
    std::vector<int> colour_coef;       // a_d
    std::vector<uint16_t> colour;       // [n_spins]
    std::vector<uint32_t> order;        // sites sorted by (colour, index)
    std::vector<uint32_t> colour_start; // [n_colours+1] into order
    // "compact" storage order (used by the multispin layout): when there are two colours and the sites
    // 2j, 2j+1 always differ in colour, site i is stored at perm[i] = colour(i)*N/2 + (i>>1), i.e. the two
    // colour classes are contiguous halves and the storage index inside a half is the RNG-SPEC colour rank.
    bool compact = false;
    std::vector<uint32_t> perm;         // [n_spins] site -> storage index (identity when !compact)
};

inline int64_t rem_euclid(int64_t a, int64_t m) {
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

// returns "" on success, else the error text
inline std::string build_plan(int n_dims, const int64_t *shape, int n_offsets, const int64_t *offsets,
                              LatticePlan &p, bool want_tables = true) {
    if (n_dims < 1 || n_dims > 8) return "n_dims must be in 1..8";
    p.n_dims = n_dims;
    p.shape.assign(shape, shape + n_dims);
    p.n_spins = 1;
    for (int d = 0; d < n_dims; d++) {
        if (shape[d] < 1) return "lattice extents must be >= 1";
        p.n_spins *= shape[d];
        if (p.n_spins >= (int64_t(1) << 32)) return "n_spins must be < 2^32";
    }
    p.strides.assign(n_dims, 1);
    for (int d = n_dims - 2; d >= 0; d--) p.strides[d] = p.strides[d + 1] * shape[d + 1];
    if (n_offsets <= 0 || offsets == nullptr) {
        p.z = n_dims;
        p.offsets.assign((size_t)n_dims * n_dims, 0);
        for (int d = 0; d < n_dims; d++) p.offsets[(size_t)d * n_dims + d] = 1;
    } else {
        p.z = n_offsets;
        p.offsets.assign(offsets, offsets + (size_t)n_offsets * n_dims);
    }
    if (p.z > 16) return "at most 16 forward neighbour directions are supported";
    p.hypercubic = (p.z == n_dims);
    for (int k = 0; k < p.z && p.hypercubic; k++)
        for (int d = 0; d < n_dims; d++)
            if (p.offsets[(size_t)k * n_dims + d] != (k == d ? 1 : 0)) p.hypercubic = false;

    // an offset that maps a site onto itself cannot be coloured (and is its own neighbour)
    for (int k = 0; k < p.z; k++) {
        bool self = true;
        for (int d = 0; d < n_dims; d++)
            if (rem_euclid(p.offsets[(size_t)k * n_dims + d], shape[d]) != 0) self = false;
        if (self) return "a neighbour offset maps every site onto itself (extent 1?): not supported by the checkerboard sweep";
    }

    // ---- colouring, step 1: linear
    p.linear_colouring = false;
    if (n_dims <= 4) {
        std::vector<int> a(n_dims);
        for (int m = 2; m <= 8 && !p.linear_colouring; m++) {
            int64_t combos = 1;
            for (int d = 0; d < n_dims; d++) combos *= m;
            for (int64_t code = 0; code < combos && !p.linear_colouring; code++) {
                int64_t c = code;
                for (int d = n_dims - 1; d >= 0; d--) { a[d] = (int)(c % m); c /= m; }  // lexicographic in (a_0, a_1, ...)
                bool ok = true;
                for (int d = 0; d < n_dims && ok; d++)
                    if ((a[d] * shape[d]) % m != 0) ok = false;
                for (int k = 0; k < p.z && ok; k++) {
                    int64_t s = 0;
                    for (int d = 0; d < n_dims; d++) s += a[d] * p.offsets[(size_t)k * n_dims + d];
                    if (rem_euclid(s, m) == 0) ok = false;
                }
                if (ok) {
                    p.linear_colouring = true;
                    p.colour_mod = m;
                    p.colour_coef = a;
                }
            }
        }
    }

    if (!want_tables && p.linear_colouring) {
        p.n_colours = p.colour_mod;
        return "";
    }

    // ---- neighbour tables (lattice.rs:66-82)
    const int z2 = 2 * p.z;
    p.nbr.assign((size_t)p.n_spins * z2, 0);
    std::vector<int64_t> coords(n_dims);
    for (int64_t i = 0; i < p.n_spins; i++) {
        for (int d = 0; d < n_dims; d++) coords[d] = (i / p.strides[d]) % shape[d];
        for (int k = 0; k < p.z; k++)
            for (int sgn = 0; sgn < 2; sgn++) {
                int64_t flat = 0;
                for (int d = 0; d < n_dims; d++) {
                    int64_t c = rem_euclid(coords[d] + (sgn ? -1 : 1) * p.offsets[(size_t)k * n_dims + d], shape[d]);
                    flat += c * p.strides[d];
                }
                p.nbr[(size_t)i * z2 + 2 * k + sgn] = (uint32_t)flat;
            }
    }

    p.colour.assign((size_t)p.n_spins, 0);
    if (p.linear_colouring) {
        p.n_colours = p.colour_mod;
        for (int64_t i = 0; i < p.n_spins; i++) {
            int64_t s = 0;
            for (int d = 0; d < n_dims; d++) s += p.colour_coef[d] * ((i / p.strides[d]) % shape[d]);
            p.colour[(size_t)i] = (uint16_t)(s % p.colour_mod);
        }
    } else {
        // ---- step 2: greedy first-fit in site order
        const uint16_t NONE = 0xFFFF;
        std::fill(p.colour.begin(), p.colour.end(), NONE);
        int nc = 0;
        std::vector<char> used(64);
        for (int64_t i = 0; i < p.n_spins; i++) {
            std::fill(used.begin(), used.end(), 0);
            for (int k = 0; k < z2; k++) {
                uint32_t j = p.nbr[(size_t)i * z2 + k];
                if (j == (uint32_t)i) return "self-neighbour: not supported by the checkerboard sweep";
                if (p.colour[j] != NONE) used[p.colour[j]] = 1;
            }
            int c = 0;
            while (used[c]) c++;
            p.colour[(size_t)i] = (uint16_t)c;
            if (c + 1 > nc) nc = c + 1;
        }
        p.n_colours = nc;
    }
    // colour-sorted order
    p.colour_start.assign((size_t)p.n_colours + 1, 0);
    for (int64_t i = 0; i < p.n_spins; i++) p.colour_start[p.colour[(size_t)i] + 1]++;
    for (int c = 0; c < p.n_colours; c++) p.colour_start[c + 1] += p.colour_start[c];
    p.order.assign((size_t)p.n_spins, 0);
    std::vector<uint32_t> fill(p.colour_start.begin(), p.colour_start.end() - 1);
    for (int64_t i = 0; i < p.n_spins; i++) p.order[fill[p.colour[(size_t)i]]++] = (uint32_t)i;
    // compact storage order
    p.compact = p.n_colours == 2 && p.n_spins % 2 == 0;
    for (int64_t i = 0; i + 1 < p.n_spins && p.compact; i += 2)
        if (p.colour[(size_t)i] == p.colour[(size_t)i + 1]) p.compact = false;
    p.perm.assign((size_t)p.n_spins, 0);
    for (int64_t i = 0; i < p.n_spins; i++)
        p.perm[(size_t)i] = p.compact ? (uint32_t)(p.colour[(size_t)i] * (p.n_spins / 2) + (i >> 1)) : (uint32_t)i;
    return "";
}

// neighbour / order tables re-expressed in storage space (identity when !compact)
inline void storage_tables(const LatticePlan &p, std::vector<uint32_t> &nbr_s, std::vector<uint32_t> &order_s) {
    const int z2 = 2 * p.z;
    nbr_s.assign(p.nbr.size(), 0);
    order_s.assign(p.order.size(), 0);
    for (int64_t i = 0; i < p.n_spins; i++) {
        const uint32_t si = p.perm[(size_t)i];
        for (int k = 0; k < z2; k++) nbr_s[(size_t)si * z2 + k] = p.perm[p.nbr[(size_t)i * z2 + k]];
    }
    for (int64_t q = 0; q < p.n_spins; q++) order_s[(size_t)q] = p.perm[p.order[(size_t)q]];
}

}  // namespace pp
