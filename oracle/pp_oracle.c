/*
 * pp_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see pp_oracle.h).
 *
 * Restates the arithmetic of PeaBrane/peapods v0.2.1 `spin-sim` for the
 * single-spin-flip sweep path.  Build with -ffp-contract=off -fno-fast-math so
 * that every f32 operation is the single IEEE operation the Rust source performs.
 * Citations are file:line under /root/reference/.
 */
#include "pp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static void set_err(const char *m) { snprintf(g_err, sizeof g_err, "%s", m); }

/* ======================================================================
 * RNG primitives
 * ====================================================================== */

/* simulation/realization.rs:9-15 and src/lib.rs:22-28 (identical bodies). */
uint64_t orc_splitmix64(uint64_t value) {
    value += 0x9E3779B97F4A7C15ull;
    uint64_t mixed = value;
    mixed = (mixed ^ (mixed >> 30)) * 0xBF58476D1CE4E5B9ull;
    mixed = (mixed ^ (mixed >> 27)) * 0x94D049BB133111EBull;
    return mixed ^ (mixed >> 31);
}

/* simulation/realization.rs:17-19 */
uint64_t orc_child_seed(uint64_t root, uint64_t domain, uint64_t index) {
    return orc_splitmix64(root ^ domain ^ orc_splitmix64(index));
}

/* src/lib.rs:30-32 */
uint64_t orc_realization_seed(uint64_t root, uint64_t r) {
    return orc_splitmix64(root ^ orc_splitmix64(r));
}

#define SYSTEM_SEED_DOMAIN 0x53A917E14C2D8B6Full /* realization.rs:6 */

/* rand_xoshiro 0.6.0 (Cargo.lock:654), Xoshiro256StarStar::seed_from_u64: the four state
 * words are successive SplitMix64 outputs (state += golden; mix).  Third-party, unpinned. */
void orc_xoshiro_seed_from_u64(uint64_t s[4], uint64_t seed) {
    uint64_t state = seed;
    for (int i = 0; i < 4; i++) {
        state += 0x9E3779B97F4A7C15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        s[i] = z ^ (z >> 31);
    }
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* xoshiro256** 1.0 (Blackman & Vigna), as implemented by rand_xoshiro 0.6.0. */
uint64_t orc_xoshiro_next_u64(uint64_t s[4]) {
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* rand_xoshiro: next_u32 takes the upper half of next_u64. */
static inline uint32_t xo_next_u32(uint64_t s[4]) { return (uint32_t)(orc_xoshiro_next_u64(s) >> 32); }

/* rand 0.8.5 Standard for f32: 24 high bits of one u32 times 2^-24 (mcmc/sweep.rs:179-181). */
static inline uint32_t xo_draw24(uint64_t s[4]) { return xo_next_u32(s) >> 8; }
static inline float u24_to_f32(uint32_t d) { return (float)d * (1.0f / 16777216.0f); }

/* rand 0.8.5 UniformInt<usize>::sample_single (widening multiply + rejection zone). */
static uint64_t xo_gen_range_usize(uint64_t s[4], uint64_t low, uint64_t high) {
    uint64_t range = high - low; /* caller guarantees high > low */
    int lz = __builtin_clzll(range);
    uint64_t zone = (range << lz) - 1;
    for (;;) {
        uint64_t v = orc_xoshiro_next_u64(s);
        __uint128_t m = (__uint128_t)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return low + hi;
    }
}

/* Philox4x32-R (Salmon et al., SC'11).  RNG-SPEC v2 uses R = ORC_PHILOX_ROUNDS = 7, the smallest round count the paper reports as
 * Crush-resistant (its Table 2; Random123 ships it as philox4x32_R(7, ...)); R = 10 is kept for the published test vectors. */
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < rounds; round++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc_philox4x32_r(ctr, key, 10, out); }
/* the RNG-SPEC generator */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc_philox4x32_r(ctr, key, ORC_PHILOX_ROUNDS, out); }

/* RNG-SPEC: 24-bit draw number `index` of the (c1, c2, c3) stream:
 * one Philox call serves four consecutive indices. */
uint32_t orc_draw24(uint64_t key, uint32_t index, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint32_t ctr[4] = {index >> 2, c1, c2, c3};
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    uint32_t out[4];
    orc_philox(ctr, k, out);
    return out[index & 3] >> 8;
}

/* ======================================================================
 * Lattice  (geometry/lattice.rs:9-109, geometry/offsets.rs:4-13)
 * ====================================================================== */
struct orc_lattice {
    int n_dims, n_neighbors;
    int64_t n_spins;
    int64_t shape[ORC_MAX_DIMS], strides[ORC_MAX_DIMS];
    uint32_t *fwd, *bwd;
};

static int64_t rem_euclid(int64_t a, int64_t m) {
    int64_t r = a % m;
    return r < 0 ? r + m : r;
}

orc_lattice *orc_lattice_new(int n_dims, const int64_t *shape, int n_offsets, const int64_t *offsets) {
    if (n_dims < 1 || n_dims > ORC_MAX_DIMS) { set_err("bad n_dims"); return NULL; }
    orc_lattice *lat = calloc(1, sizeof *lat);
    lat->n_dims = n_dims;
    int64_t *hyper = NULL;
    if (!offsets) { /* geometry/offsets.rs:4-13 */
        n_offsets = n_dims;
        hyper = calloc((size_t)n_dims * n_dims, sizeof(int64_t));
        for (int d = 0; d < n_dims; d++) hyper[d * n_dims + d] = 1;
        offsets = hyper;
    }
    lat->n_neighbors = n_offsets;
    lat->n_spins = 1;
    for (int d = 0; d < n_dims; d++) { lat->shape[d] = shape[d]; lat->n_spins *= shape[d]; }
    /* lattice.rs:58-61 row-major strides */
    for (int d = 0; d < n_dims; d++) lat->strides[d] = 1;
    for (int d = n_dims - 2; d >= 0; d--) lat->strides[d] = lat->strides[d + 1] * shape[d + 1];
    size_t tab = (size_t)lat->n_spins * n_offsets;
    lat->fwd = malloc(tab * sizeof(uint32_t) + 4);
    lat->bwd = malloc(tab * sizeof(uint32_t) + 4);
    /* lattice.rs:66-82 */
    for (int64_t i = 0; i < lat->n_spins; i++) {
        int64_t coords[ORC_MAX_DIMS];
        for (int d = 0; d < n_dims; d++) coords[d] = (i / lat->strides[d]) % lat->shape[d];
        for (int k = 0; k < n_offsets; k++) {
            for (int sgn = 0; sgn < 2; sgn++) {
                int64_t sign = sgn == 0 ? 1 : -1;
                int64_t flat = 0;
                for (int d = 0; d < n_dims; d++) {
                    int64_t c = rem_euclid(coords[d] + sign * offsets[k * n_dims + d], lat->shape[d]);
                    flat += c * lat->strides[d];
                }
                (sgn == 0 ? lat->fwd : lat->bwd)[i * n_offsets + k] = (uint32_t)flat;
            }
        }
    }
    free(hyper);
    return lat;
}

void orc_lattice_free(orc_lattice *lat) {
    if (!lat) return;
    free(lat->fwd); free(lat->bwd); free(lat);
}
int64_t orc_lattice_n_spins(const orc_lattice *lat) { return lat->n_spins; }
int orc_lattice_n_neighbors(const orc_lattice *lat) { return lat->n_neighbors; }
int64_t orc_lattice_stride(const orc_lattice *lat, int d) { return lat->strides[d]; }
uint32_t orc_neighbor_fwd(const orc_lattice *lat, int64_t i, int d) { return lat->fwd[i * lat->n_neighbors + d]; }
uint32_t orc_neighbor_bwd(const orc_lattice *lat, int64_t i, int d) { return lat->bwd[i * lat->n_neighbors + d]; }

int orc_colouring_is_valid(const orc_lattice *lat, const uint16_t *colour) {
    int z = lat->n_neighbors;
    for (int64_t i = 0; i < lat->n_spins; i++)
        for (int d = 0; d < z; d++) {
            if (colour[lat->fwd[i * z + d]] == colour[i]) return 0;
            if (colour[lat->bwd[i * z + d]] == colour[i]) return 0;
        }
    return 1;
}

/* ======================================================================
 * Acceptance rules  (mcmc/sweep.rs:8-19, 35-48, 99-185, 247-257, 271-283)
 * ====================================================================== */
#define F32_UNIFORM_VALUES (1u << 24) /* sweep.rs:99-100 */

/* sweep.rs:161-165 */
int orc_metropolis_legacy_accepts(float temperature, int32_t energy_change, uint32_t draw) {
    float uniform = (float)draw / (float)F32_UNIFORM_VALUES;
    return (float)energy_change >= (temperature / 2.0f) * logf(uniform);
}

/* sweep.rs:147-159 */
uint32_t orc_metropolis_accepted_count(float temperature, int32_t energy_change) {
    uint32_t low = 0, high = F32_UNIFORM_VALUES;
    while (low < high) {
        uint32_t mid = low + (high - low) / 2;
        if (orc_metropolis_legacy_accepts(temperature, energy_change, mid)) low = mid + 1;
        else high = mid;
    }
    return low;
}

/* heat-bath rule of sweep.rs:279-282 on the same 24-bit grid */
int orc_gibbs_legacy_accepts(float temperature, int32_t energy_change, uint32_t draw) {
    float u = (float)draw / (float)F32_UNIFORM_VALUES;
    return (float)energy_change >= (temperature / 2.0f) * logf(u / (1.0f - u));
}

uint32_t orc_gibbs_accepted_count(float temperature, int32_t energy_change) {
    uint32_t low = 0, high = F32_UNIFORM_VALUES;
    while (low < high) {
        uint32_t mid = low + (high - low) / 2;
        if (orc_gibbs_legacy_accepts(temperature, energy_change, mid)) low = mid + 1;
        else high = mid;
    }
    return low;
}

/* sweep.rs:108-145: eligibility gate + table build; -1 = None (fail closed) */
int orc_metropolis_lookup(const float *couplings, int64_t n_couplings, const float *temps, int n_temps,
                          int n_neighbors, uint32_t *table) {
    for (int64_t i = 0; i < n_couplings; i++) {
        float c = couplings[i];
        if (!(c == -1.0f || c == 0.0f || c == 1.0f)) return -1;
    }
    for (int t = 0; t < n_temps; t++)
        if (!(isfinite(temps[t]) && temps[t] / 2.0f > 0.0f)) return -1;
    int offset = 2 * n_neighbors;
    int width = 2 * offset + 1;
    if (table)
        for (int t = 0; t < n_temps; t++)
            for (int ec = -offset; ec <= offset; ec++)
                table[t * width + ec + offset] = orc_metropolis_accepted_count(temps[t], ec);
    return 0;
}

/* sweep.rs:8-19 */
static inline float local_field(const orc_lattice *lat, const int8_t *s, const float *J, int64_t i) {
    int z = lat->n_neighbors;
    float h = 0.0f;
    for (int d = 0; d < z; d++) {
        uint32_t jf = lat->fwd[i * z + d];
        h += (float)s[jf] * J[i * z + d];
        uint32_t jb = lat->bwd[i * z + d];
        h += (float)s[jb] * J[(int64_t)jb * z + d];
    }
    return h;
}

/* One attempt given the 24-bit draw.  lookup != NULL: sweep.rs:170-185; else the log forms
 * sweep.rs:35-48 with thresholds sweep.rs:256 (Metropolis) / sweep.rs:279-282 (Gibbs). */
static inline void attempt(int8_t *s, int64_t i, float h, uint32_t draw, float temp, int sweep_mode,
                           const uint32_t *lookup_row, int offset) {
    if (lookup_row) {
        int32_t ec = (int32_t)((float)(-s[i]) * h);
        if (draw < lookup_row[ec + offset]) s[i] = -s[i];
        return;
    }
    float si = (float)s[i];
    float eng_change = -si * h;
    float u = u24_to_f32(draw);
    float thr = sweep_mode == ORC_SWEEP_METROPOLIS ? (temp / 2.0f) * logf(u)
                                                   : (temp / 2.0f) * logf(u / (1.0f - u));
    if (eng_change >= thr) s[i] = -s[i];
}

/* mcmc/sweep.rs:220-284 through parallel.rs:27-33, typewriter order (sweep.rs:51-97; the
 * canonical-square fast path is bit-identical to the table path, sweep.rs:293-343). */
void orc_sweep_xoshiro(const orc_lattice *lat, int8_t *spins, const float *couplings,
                       const float *temperatures, const int64_t *system_ids, int64_t n_systems,
                       uint64_t *rng_states, int sweep_mode, int use_lookup) {
    int z = lat->n_neighbors, offset = 2 * z, width = 4 * z + 1;
    uint32_t *table = NULL;
    if (use_lookup && sweep_mode == ORC_SWEEP_METROPOLIS) {
        table = malloc(sizeof(uint32_t) * (size_t)n_systems * width);
        if (orc_metropolis_lookup(couplings, lat->n_spins * z, temperatures, (int)n_systems, z, table) != 0) {
            free(table); table = NULL;
        }
    }
    for (int64_t slot = 0; slot < n_systems; slot++) {
        int64_t sys = system_ids[slot];
        int8_t *s = spins + sys * lat->n_spins;
        uint64_t *rng = rng_states + sys * 4;
        float temp = temperatures[slot];
        const uint32_t *row = table ? table + slot * width : NULL;
        for (int64_t i = 0; i < lat->n_spins; i++) {
            float h = local_field(lat, s, couplings, i);
            attempt(s, i, h, xo_draw24(rng), temp, sweep_mode, row, offset);
        }
    }
    free(table);
}

/* colour-ordered site list + rank-within-colour (RNG-SPEC) */
static void colour_order(const orc_lattice *lat, const uint16_t *colour, int64_t **order_out,
                         uint32_t **rank_out, int *n_colours_out) {
    int64_t n = lat->n_spins;
    int nc = 0;
    for (int64_t i = 0; i < n; i++) if (colour[i] + 1 > nc) nc = colour[i] + 1;
    int64_t *start = calloc((size_t)nc + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) start[colour[i] + 1]++;
    for (int c = 0; c < nc; c++) start[c + 1] += start[c];
    int64_t *order = malloc(sizeof(int64_t) * (size_t)n);
    uint32_t *rank = malloc(sizeof(uint32_t) * (size_t)n);
    int64_t *fill = calloc((size_t)nc, sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        int c = colour[i];
        rank[i] = (uint32_t)fill[c];
        order[start[c] + fill[c]++] = i;
    }
    free(start); free(fill);
    *order_out = order; *rank_out = rank; *n_colours_out = nc;
}

/* RNG-SPEC packed mapping (bit-packed single-lattice kernels, pp_kernels_slabp.cuh): the 32 ranks of a block r >> 5 share
 * SIX generator calls, counter = {r >> 5, sweep, stream, tag | call << 8 | colour}, whose 24 output words W[0..23] are cut into 32
 * fields of 24 bits: for g = 0..7 and A, B, C = W[3g], W[3g+1], W[3g+2] the ranks 4g .. 4g+3 of the block draw
 * A >> 8, B >> 8, C >> 8 and (A & 255) << 16 | (B & 255) << 8 | (C & 255).  Every generated bit is used once. */
uint32_t orc_draw24_packed(uint64_t key, uint32_t rank, uint32_t sweep, uint32_t stream, uint32_t tag_colour) {
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)}, w[24];
    for (uint32_t call = 0; call < 6; call++) {
        uint32_t ctr[4] = {rank >> 5, sweep, stream, tag_colour | (call << 8)};
        orc_philox(ctr, k, w + 4 * call);
    }
    uint32_t b = rank & 31u, g = b >> 2, j = b & 3u;
    const uint32_t *t = w + 3 * g;
    if (j < 3) return t[j] >> 8;
    return ((t[0] & 255u) << 16) | ((t[1] & 255u) << 8) | (t[2] & 255u);
}

static void sweep_philox_impl(const orc_lattice *lat, int8_t *spins, const float *couplings,
                              const float *temperatures, const int64_t *system_ids, int64_t n_systems,
                              const uint16_t *colour, const int64_t *order, const uint32_t *rank,
                              uint64_t key, uint32_t sweep_index, int sweep_mode,
                              const uint32_t *table, int stream_is_slot, int packed) {
    int z = lat->n_neighbors, offset = 2 * z, width = 4 * z + 1;
    int sysq = packed == 2; /* system-quad mapping (pp_kernels_swords.cuh): one call per (site rank, four consecutive systems) */
    if (sysq) packed = 0;
    uint32_t tag = sysq ? ORC_TAG_SWEEP_SYSQ : packed ? ORC_TAG_SWEEP_PACKED : stream_is_slot ? ORC_TAG_SWEEP_MSC : ORC_TAG_SWEEP;
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    for (int64_t slot = 0; slot < n_systems; slot++) {
        int64_t sys = system_ids[slot];
        int8_t *s = spins + sys * lat->n_spins;
        float temp = temperatures[slot];
        const uint32_t *row = table ? table + slot * width : NULL;
        uint32_t stream = (uint32_t)(stream_is_slot ? slot : sys);
        uint32_t cached_c0 = 0xFFFFFFFFu, cached_c3 = 0, out[24] = {0};
        for (int64_t p = 0; p < lat->n_spins; p++) {
            int64_t i = order[p];
            uint32_t c3 = tag | colour[i];
            uint32_t r = rank[i];
            uint32_t draw;
            if (sysq) {
                uint32_t ctr[4] = {r, sweep_index, stream >> 2, c3};
                orc_philox(ctr, k, out);
                draw = out[stream & 3u] >> 8;
            } else if (packed) {
                if ((r >> 5) != cached_c0 || c3 != cached_c3) {
                    for (uint32_t call = 0; call < 6; call++) {
                        uint32_t ctr[4] = {r >> 5, sweep_index, stream, c3 | (call << 8)};
                        orc_philox(ctr, k, out + 4 * call);
                    }
                    cached_c0 = r >> 5; cached_c3 = c3;
                }
                const uint32_t *t = out + 3 * ((r & 31u) >> 2);
                draw = (r & 3u) < 3u ? t[r & 3u] >> 8 : ((t[0] & 255u) << 16) | ((t[1] & 255u) << 8) | (t[2] & 255u);
            } else {
                if ((r >> 2) != cached_c0 || c3 != cached_c3) {
                    uint32_t ctr[4] = {r >> 2, sweep_index, stream, c3};
                    orc_philox(ctr, k, out);
                    cached_c0 = r >> 2; cached_c3 = c3;
                }
                draw = out[r & 3] >> 8;
            }
            float h = local_field(lat, s, couplings, i);
            attempt(s, i, h, draw, temp, sweep_mode, row, offset);
        }
    }
}

/* +-J Gibbs uses the integer heat-bath table (same rule as the log form on the 24-bit grid) */
static uint32_t *build_table(const float *couplings, int64_t n_couplings, const float *temps,
                             int64_t n_systems, int z, int sweep_mode) {
    int width = 4 * z + 1, offset = 2 * z;
    if (orc_metropolis_lookup(couplings, n_couplings, temps, (int)n_systems, z, NULL) != 0) return NULL;
    uint32_t *table = malloc(sizeof(uint32_t) * (size_t)n_systems * width);
    for (int64_t t = 0; t < n_systems; t++)
        for (int ec = -offset; ec <= offset; ec++)
            table[t * width + ec + offset] = sweep_mode == ORC_SWEEP_METROPOLIS
                                                 ? orc_metropolis_accepted_count(temps[t], ec)
                                                 : orc_gibbs_accepted_count(temps[t], ec);
    return table;
}

void orc_sweep_philox(const orc_lattice *lat, int8_t *spins, const float *couplings,
                      const float *temperatures, const int64_t *system_ids, int64_t n_systems,
                      const uint16_t *colour, uint64_t key, uint32_t sweep_index, int sweep_mode,
                      int use_lookup, int stream_is_slot) {
    int64_t *order; uint32_t *rank; int nc;
    colour_order(lat, colour, &order, &rank, &nc);
    uint32_t *table = use_lookup ? build_table(couplings, lat->n_spins * lat->n_neighbors, temperatures,
                                               n_systems, lat->n_neighbors, sweep_mode)
                                 : NULL;
    sweep_philox_impl(lat, spins, couplings, temperatures, system_ids, n_systems, colour, order, rank,
                      key, sweep_index, sweep_mode, table, stream_is_slot & 1, (stream_is_slot & 4) ? 2 : (stream_is_slot >> 1) & 1);
    free(table); free(order); free(rank);
}

/* ======================================================================
 * Energy / magnetisation  (spins/energy.rs:78-110)
 * ====================================================================== */
void orc_energies_mags(const orc_lattice *lat, const int8_t *spins, const float *couplings,
                       int64_t n_systems, float *energies, int64_t *mags) {
    int64_t n = lat->n_spins;
    int z = lat->n_neighbors;
    for (int64_t r = 0; r < n_systems; r++) {
        const int8_t *s = spins + r * n;
        float total = 0.0f;
        int64_t m = 0;
        for (int64_t i = 0; i < n; i++) {
            int8_t spin = s[i];
            m += spin;
            float si = (float)spin;
            for (int d = 0; d < z; d++) {
                uint32_t j = lat->fwd[i * z + d];
                float sj = (float)s[j];
                float c = couplings[i * z + d];
                float interaction = si * sj * c;
                total += interaction;
            }
        }
        energies[r] = total / (float)n;
        if (mags) mags[r] = m;
    }
}

/* statistics/overlap.rs:259-281 */
void orc_overlap_dots(const orc_lattice *lat, const int8_t *a, const int8_t *b, int64_t *dot_spin,
                      int64_t *dot_link) {
    int64_t ds = 0, dl = 0;
    int z = lat->n_neighbors;
    for (int64_t j = 0; j < lat->n_spins; j++) {
        int64_t q = (int64_t)a[j] * (int64_t)b[j];
        ds += q;
        for (int d = 0; d < z; d++) {
            uint32_t k = lat->fwd[j * z + d];
            int64_t nq = (int64_t)a[k] * (int64_t)b[k];
            dl += q * nq;
        }
    }
    *dot_spin = ds; *dot_link = dl;
}

/* mcmc/tempering.rs:59-70 */
int orc_full_ladder_edges(int n_temps, int first_parity, int32_t *edges_out) {
    int n = 0;
    if (n_temps < 2) return 0;
    int parities[2] = {first_parity, 1 - first_parity};
    for (int pi = 0; pi < 2; pi++)
        for (int edge = parities[pi]; edge < n_temps - 1; edge += 2) edges_out[n++] = edge;
    return n;
}

/* ======================================================================
 * Realization state  (simulation/realization.rs:21-246)
 * ====================================================================== */
typedef struct {
    uint64_t *edge_attempts, *edge_acceptances, *round_trips;
    uint8_t *trip_state;
    int next_parity, cold_slot, hot_slot;
} pt_state;

typedef struct {
    const float *couplings; /* [N*z'] borrowed from sim->couplings */
    int8_t *spins;          /* [S*N] */
    float *temperatures;    /* [S] = temps repeated R times (realization.rs:166) */
    int64_t *system_ids;    /* [S] slot -> system */
    uint64_t *rngs;         /* [S][4] xoshiro states */
    float *energies;        /* [S] by system */
    pt_state pt;
    uint64_t base_seed;
} realization;

struct orc_sim {
    orc_lattice *lat;
    int64_t n_real;
    int n_replicas, n_temps, rng_mode;
    float *couplings; /* [D][N*z'] */
    float *temps;     /* [T] */
    uint64_t ctor_seed, cur_seed;
    int64_t sample_offset; /* global index of realization 0 (a shard of a larger run); seeds use global indices */
    realization *reals;
    uint16_t *colour;
    int64_t *order;
    uint32_t *rank;
    int n_colours;
    uint32_t sweep_counter, pt_event_counter; /* RNG-SPEC counters; persist across sample() */
};

/* realization.rs:59-67, 92-107 */
static void pt_reset(pt_state *pt, int n_replicas, int n_temps, const int64_t *system_ids,
                     const float *temperatures) {
    int n_edges = n_temps > 0 ? n_temps - 1 : 0;
    memset(pt->edge_attempts, 0, sizeof(uint64_t) * (size_t)n_edges);
    memset(pt->edge_acceptances, 0, sizeof(uint64_t) * (size_t)n_edges);
    memset(pt->round_trips, 0, sizeof(uint64_t) * (size_t)n_replicas * n_temps);
    memset(pt->trip_state, 0, (size_t)n_replicas * n_temps);
    pt->next_parity = 0;
    pt->cold_slot = pt->hot_slot = 0;
    for (int slot = 1; slot < n_temps; slot++) {
        if (temperatures[slot] < temperatures[pt->cold_slot]) pt->cold_slot = slot;
        if (temperatures[slot] > temperatures[pt->hot_slot]) pt->hot_slot = slot;
    }
    if (n_temps == 0) return;
    for (int r = 0; r < n_replicas; r++) pt->trip_state[system_ids[r * n_temps + pt->hot_slot]] = 1;
}

/* realization.rs:109-120 */
static void pt_record_arrival(pt_state *pt, int64_t system, int slot) {
    if (slot == pt->hot_slot) {
        if (pt->trip_state[system] == 2) pt->round_trips[system] += 1;
        pt->trip_state[system] = 1;
        return;
    }
    if (slot == pt->cold_slot && pt->trip_state[system] == 1) pt->trip_state[system] = 2;
}

/* realization.rs:73-82 */
static void pt_record_attempt(pt_state *pt, int edge, int accepted, int64_t left, int64_t right) {
    pt->edge_attempts[edge] += 1;
    if (!accepted) return;
    pt->edge_acceptances[edge] += 1;
    pt_record_arrival(pt, left, edge + 1);
    pt_record_arrival(pt, right, edge);
}

/* Replays a list of attempts through PtState exactly as realization.rs:73-120 does
 * (used by the hot->cold->hot KAT, realization.rs:285-302). */
void orc_pt_replay(int n_replicas, int n_temps, const float *temps, int n_attempts, const int32_t *edges,
                   const int32_t *accepted, const int64_t *left, const int64_t *right,
                   uint64_t *edge_attempts, uint64_t *edge_acceptances, uint64_t *round_trips) {
    int S = n_replicas * n_temps;
    pt_state pt;
    pt.edge_attempts = edge_attempts; pt.edge_acceptances = edge_acceptances; pt.round_trips = round_trips;
    pt.trip_state = calloc((size_t)S + 1, 1);
    int64_t *ids = malloc(sizeof(int64_t) * (size_t)(S + 1));
    float *tt = malloc(sizeof(float) * (size_t)(S + 1));
    for (int i = 0; i < S; i++) { ids[i] = i; tt[i] = temps[i % n_temps]; }
    pt_reset(&pt, n_replicas, n_temps, ids, tt);
    for (int a = 0; a < n_attempts; a++) pt_record_attempt(&pt, edges[a], accepted[a], left[a], right[a]);
    free(pt.trip_state); free(ids); free(tt);
}

/* realization.rs:166-207 (new) and :213-246 (reset): seed, draw spins, identity ids, energies */
static void realization_init(orc_sim *sim, realization *re, uint64_t base_seed) {
    const orc_lattice *lat = sim->lat;
    int64_t n = lat->n_spins;
    int S = sim->n_replicas * sim->n_temps;
    re->base_seed = base_seed;
    for (int i = 0; i < S; i++) {
        if (sim->rng_mode == ORC_RNG_XOSHIRO) {
            uint64_t *rng = re->rngs + 4 * i;
            orc_xoshiro_seed_from_u64(rng, orc_child_seed(base_seed, SYSTEM_SEED_DOMAIN, (uint64_t)i));
            /* realization.rs:177-182: gen::<f32>() < 0.5 -> -1 */
            for (int64_t j = 0; j < n; j++)
                re->spins[i * n + j] = u24_to_f32(xo_draw24(rng)) < 0.5f ? -1 : 1;
        } else {
            /* RNG-SPEC INIT: same rule on the Philox draw of (system, site) */
            for (int64_t j = 0; j < n; j++)
                re->spins[i * n + j] =
                    u24_to_f32(orc_draw24(base_seed, (uint32_t)j, 0u, (uint32_t)i, ORC_TAG_INIT)) < 0.5f ? -1 : 1;
        }
    }
    for (int i = 0; i < S; i++) re->system_ids[i] = i;
    orc_energies_mags(lat, re->spins, re->couplings, S, re->energies, NULL);
    pt_reset(&re->pt, sim->n_replicas, sim->n_temps, re->system_ids, re->temperatures);
}

orc_sim *orc_sim_new(int n_dims, const int64_t *shape, int n_offsets, const int64_t *offsets,
                     const float *couplings, int64_t n_real, const float *temps, int n_temps,
                     int n_replicas, uint64_t seed, int rng_mode, const uint16_t *colour) {
    orc_lattice *lat = orc_lattice_new(n_dims, shape, n_offsets, offsets);
    if (!lat) return NULL;
    if (rng_mode != ORC_RNG_XOSHIRO) {
        if (!colour || !orc_colouring_is_valid(lat, colour)) {
            set_err("PHILOX modes need a valid colouring");
            orc_lattice_free(lat);
            return NULL;
        }
    }
    orc_sim *sim = calloc(1, sizeof *sim);
    sim->lat = lat;
    sim->n_real = n_real; sim->n_replicas = n_replicas; sim->n_temps = n_temps; sim->rng_mode = rng_mode;
    int64_t n = lat->n_spins;
    int z = lat->n_neighbors;
    int S = n_replicas * n_temps;
    size_t chunk = (size_t)n * z;
    sim->couplings = malloc(sizeof(float) * chunk * (size_t)n_real);
    memcpy(sim->couplings, couplings, sizeof(float) * chunk * (size_t)n_real);
    sim->temps = malloc(sizeof(float) * (size_t)(n_temps > 0 ? n_temps : 1));
    memcpy(sim->temps, temps, sizeof(float) * (size_t)n_temps);
    sim->ctor_seed = sim->cur_seed = seed;
    if (colour) {
        sim->colour = malloc(sizeof(uint16_t) * (size_t)n);
        memcpy(sim->colour, colour, sizeof(uint16_t) * (size_t)n);
        colour_order(lat, sim->colour, &sim->order, &sim->rank, &sim->n_colours);
    }
    sim->reals = calloc((size_t)n_real, sizeof(realization));
    for (int64_t r = 0; r < n_real; r++) {
        realization *re = &sim->reals[r];
        re->couplings = sim->couplings + chunk * (size_t)r;
        re->spins = malloc((size_t)S * n);
        re->temperatures = malloc(sizeof(float) * (size_t)(S > 0 ? S : 1));
        for (int k = 0; k < S; k++) re->temperatures[k] = temps[k % n_temps];
        re->system_ids = malloc(sizeof(int64_t) * (size_t)(S > 0 ? S : 1));
        re->rngs = calloc((size_t)(S > 0 ? S : 1) * 4, sizeof(uint64_t));
        re->energies = calloc((size_t)(S > 0 ? S : 1), sizeof(float));
        int n_edges = n_temps > 0 ? n_temps - 1 : 0;
        re->pt.edge_attempts = calloc((size_t)n_edges + 1, sizeof(uint64_t));
        re->pt.edge_acceptances = calloc((size_t)n_edges + 1, sizeof(uint64_t));
        re->pt.round_trips = calloc((size_t)S + 1, sizeof(uint64_t));
        re->pt.trip_state = calloc((size_t)S + 1, 1);
        realization_init(sim, re, orc_realization_seed(seed, (uint64_t)(sim->sample_offset + r))); /* lib.rs:158-164 */
    }
    return sim;
}

void orc_sim_free(orc_sim *sim) {
    if (!sim) return;
    for (int64_t r = 0; r < sim->n_real; r++) {
        realization *re = &sim->reals[r];
        free(re->spins); free(re->temperatures); free(re->system_ids); free(re->rngs); free(re->energies);
        free(re->pt.edge_attempts); free(re->pt.edge_acceptances); free(re->pt.round_trips);
        free(re->pt.trip_state);
    }
    free(sim->reals); free(sim->couplings); free(sim->temps); free(sim->colour); free(sim->order);
    free(sim->rank);
    orc_lattice_free(sim->lat);
    free(sim);
}

/* src/lib.rs:624-633 */
void orc_sim_reset(orc_sim *sim, int has_seed, uint64_t seed) {
    uint64_t base = has_seed ? seed : sim->ctor_seed;
    sim->cur_seed = base;
    sim->sweep_counter = 0;
    sim->pt_event_counter = 0;
    for (int64_t r = 0; r < sim->n_real; r++)
        realization_init(sim, &sim->reals[r], orc_realization_seed(base, (uint64_t)(sim->sample_offset + r)));
}

/* shard support: this sim holds realizations [offset, offset + D) of a larger run; re-initialises the state */
void orc_sim_set_sample_offset(orc_sim *sim, int64_t offset) {
    sim->sample_offset = offset;
    orc_sim_reset(sim, 1, sim->cur_seed);
}

const int8_t *orc_sim_spins(const orc_sim *sim, int64_t r) { return sim->reals[r].spins; }
const int64_t *orc_sim_system_ids(const orc_sim *sim, int64_t r) { return sim->reals[r].system_ids; }
const float *orc_sim_energies(const orc_sim *sim, int64_t r) { return sim->reals[r].energies; }

/* ======================================================================
 * Sweep loop for one realization  (simulation/mod.rs:177-863, hot-path lines only)
 * ====================================================================== */
/* ---- clusters/fk.rs:28-171 (union-find path), RNG-SPEC draws ------------------------------------ */
uint32_t orc_fk_bond_count(float temperature) {
    float p = 1.0f - expf(-2.0f * 1.0f / temperature); /* fk.rs:113 with interaction = 1 */
    if (!(p > 0.0f)) return 0;
    double c = ceil((double)p * 16777216.0);             /* u = draw * 2^-24 is exact: u < p <=> draw < ceil(p * 2^24) */
    if (c > 16777216.0) c = 16777216.0;
    return (uint32_t)c;
}
static int64_t uf_find(int64_t *parent, int64_t i) {
    while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
    return i;
}
void orc_fk_update(const orc_lattice *lat, int8_t *spins, const float *couplings, float temperature, uint64_t key,
                   uint32_t sweep_index, uint32_t system_id, int wolff) {
    int64_t N = lat->n_spins;
    int z = lat->n_neighbors;
    uint32_t count = orc_fk_bond_count(temperature);
    int64_t *parent = malloc(sizeof(int64_t) * (size_t)N);
    for (int64_t i = 0; i < N; i++) parent[i] = i;
    for (int64_t i = 0; i < N; i++)
        for (int d = 0; d < z; d++) { /* fk.rs:107-115: forward bonds, each once */
            int64_t j = lat->fwd[i * z + d];
            float inter = (float)spins[i] * (float)spins[j] * couplings[i * z + d];
            if (inter <= 0.0f) continue;
            if (orc_draw24(key, (uint32_t)(i * z + d), sweep_index, system_id, ORC_TAG_FK_BOND) >= count) continue;
            int64_t a = uf_find(parent, i), b = uf_find(parent, j);
            if (a < b) parent[b] = a; else if (b < a) parent[a] = b; /* the smallest site labels the cluster */
        }
    for (int64_t i = 0; i < N; i++) parent[i] = uf_find(parent, i);
    if (wolff) { /* fk.rs:151-158 */
        uint32_t ctr[4] = {0xFFFFFFFFu, sweep_index, system_id, ORC_TAG_FK_FLIP}, k[2] = {(uint32_t)key, (uint32_t)(key >> 32)}, o[4];
        orc_philox(ctr, k, o);
        int64_t seed = (int64_t)(((uint64_t)o[1] * (uint64_t)N) >> 32);
        int64_t root = parent[seed];
        for (int64_t i = 0; i < N; i++) if (parent[i] == root) spins[i] = (int8_t)-spins[i];
    } else { /* fk.rs:159-170: one fair draw per cluster */
        for (int64_t i = 0; i < N; i++)
            if (orc_draw24(key, (uint32_t)parent[i], sweep_index, system_id, ORC_TAG_FK_FLIP) < (1u << 23)) spins[i] = (int8_t)-spins[i];
    }
    free(parent);
}

/* ---- clusters/overlap.rs:34-56, 146-339 (Houdayer, group size 2), RNG-SPEC draws ------------------ */
void orc_houdayer_slot(const orc_lattice *lat, int8_t *spins, const int64_t *system_ids, int n_temps, int n_replicas, int t,
                       uint64_t key, uint32_t sweep_index, int wolff) {
    int64_t N = lat->n_spins;
    int z = lat->n_neighbors, R = n_replicas, P = R / 2;
    int64_t sys[64];
    for (int k = 0; k < R; k++) sys[k] = system_ids[k * n_temps + t]; /* overlap.rs:45-48 */
    for (int i = R - 1; i >= 1; i--) { /* shuffle (overlap.rs:49): Fisher-Yates, step i draws j in [0, i] */
        uint32_t ctr[4] = {(uint32_t)i, sweep_index, (uint32_t)t, ORC_TAG_OC_PAIR}, k2[2] = {(uint32_t)key, (uint32_t)(key >> 32)}, o[4];
        orc_philox(ctr, k2, o);
        int j = (int)(((uint64_t)o[0] * (uint64_t)(i + 1)) >> 32);
        int64_t tmp = sys[i]; sys[i] = sys[j]; sys[j] = tmp;
    }
    int64_t *parent = malloc(sizeof(int64_t) * (size_t)N);
    uint8_t *active = malloc((size_t)N), *multi = malloc((size_t)N);
    for (int g = 0; g < P; g++) {
        int8_t *a = spins + sys[2 * g] * N, *b = spins + sys[2 * g + 1] * N;
        uint32_t stream = (uint32_t)(t * P + g);
        for (int64_t i = 0; i < N; i++) { parent[i] = i; active[i] = a[i] != b[i]; multi[i] = 0; } /* overlap.rs:222-228 */
        for (int64_t i = 0; i < N; i++)
            for (int d = 0; d < z; d++) { /* overlap.rs:231-234: deterministic bonds between active neighbours */
                int64_t j = lat->fwd[i * z + d];
                if (!(active[i] && active[j]) || i == j) continue;
                multi[i] = multi[j] = 1;
                int64_t ra = uf_find(parent, i), rb = uf_find(parent, j);
                if (ra < rb) parent[rb] = ra; else if (rb < ra) parent[ra] = rb;
            }
        for (int64_t i = 0; i < N; i++) parent[i] = uf_find(parent, i);
        if (wolff) { /* overlap.rs:245-256 */
            uint64_t best = UINT64_MAX;
            for (int64_t i = 0; i < N; i++)
                if (active[i]) {
                    uint64_t score = ((uint64_t)orc_draw24(key, (uint32_t)i, sweep_index, stream, ORC_TAG_OC_SEED) << 32) | (uint64_t)i;
                    if (score < best) best = score;
                }
            if (best == UINT64_MAX) continue; /* no active site: nothing to do */
            int64_t root = parent[best & 0xFFFFFFFFu];
            for (int64_t i = 0; i < N; i++)
                if (active[i] && parent[i] == root) { a[i] = (int8_t)-a[i]; b[i] = (int8_t)-b[i]; }
        } else { /* overlap.rs:293-307: clusters of more than one site, fair coin each */
            for (int64_t i = 0; i < N; i++)
                if (active[i] && multi[i] &&
                    orc_draw24(key, (uint32_t)parent[i], sweep_index, stream, ORC_TAG_OC_FLIP) < (1u << 23)) {
                    a[i] = (int8_t)-a[i]; b[i] = (int8_t)-b[i];
                }
        }
    }
    free(parent); free(active); free(multi);
}

/* ---- statistics/autocorrelation.rs: ring backend ------------------------------------------------ */
typedef struct {
    int max_lag, n_temps, ring_len, ring_pos;
    int64_t n_recorded;
    double *sum_o, *sum_o2, *sum_prod; /* [T], [T], [T][max_lag+1] */
    float *ring;                       /* [T][ring_len] */
} autocorr;

static autocorr *ac_new(int max_lag, int n_temps) { /* :33-65 */
    autocorr *a = calloc(1, sizeof *a);
    a->max_lag = max_lag; a->n_temps = n_temps; a->ring_len = max_lag + 1;
    a->sum_o = calloc((size_t)n_temps, sizeof(double));
    a->sum_o2 = calloc((size_t)n_temps, sizeof(double));
    a->sum_prod = calloc((size_t)n_temps * (size_t)(max_lag + 1), sizeof(double));
    a->ring = calloc((size_t)n_temps * (size_t)(max_lag + 1), sizeof(float));
    return a;
}
static void ac_free(autocorr *a) {
    if (!a) return;
    free(a->sum_o); free(a->sum_o2); free(a->sum_prod); free(a->ring); free(a);
}
static void ac_push(autocorr *a, const double *values) { /* :68-112 */
    for (int t = 0; t < a->n_temps; t++) {
        float o = (float)values[t];
        a->sum_o[t] += (double)o;
        a->sum_o2[t] += (double)o * (double)o;
    }
    int pos = a->ring_pos;
    int64_t n_back = a->n_recorded < a->max_lag ? a->n_recorded : a->max_lag;
    for (int t = 0; t < a->n_temps; t++) {
        float o = (float)values[t];
        float *ring = a->ring + (size_t)t * a->ring_len;
        double *sp = a->sum_prod + (size_t)t * (a->max_lag + 1);
        ring[pos] = o;
        int64_t no_wrap = pos < n_back ? pos : n_back;
        for (int64_t delta = 0; delta <= no_wrap; delta++) sp[delta] += (double)o * (double)ring[pos - delta];
        for (int64_t delta = pos + 1; delta <= n_back; delta++) sp[delta] += (double)o * (double)ring[pos + a->ring_len - delta];
    }
    a->ring_pos = (pos + 1) % a->ring_len;
    a->n_recorded++;
}
static void ac_finish(const autocorr *a, double *gamma) { /* :114-124, :166-199; gamma[T][max_lag+1] */
    int L1 = a->max_lag + 1;
    for (int t = 0; t < a->n_temps; t++) {
        double *g = gamma + (size_t)t * L1;
        int degenerate = a->n_recorded == 0;
        double mean = 0.0, var = 0.0, m = (double)a->n_recorded;
        if (!degenerate) {
            mean = a->sum_o[t] / m;
            var = a->sum_o2[t] / m - mean * mean;
            if (var <= 0.0) degenerate = 1;
        }
        for (int delta = 0; delta < L1; delta++) {
            if (degenerate) { g[delta] = delta == 0 ? 1.0 : 0.0; continue; }
            int64_t cnt = a->n_recorded > delta ? a->n_recorded - delta : 0;
            if (cnt <= 0) { g[delta] = delta == 0 ? 1.0 : 0.0; continue; }
            g[delta] = (a->sum_prod[(size_t)t * L1 + delta] / (double)cnt - mean * mean) / var;
        }
    }
}
double orc_sokal_tau(const double *gamma, int n) { /* :201-210 */
    double tau = 0.5;
    for (int w = 1; w < n; w++) {
        tau += gamma[w];
        if ((double)w >= 5.0 * tau) return tau;
    }
    return tau;
}
void orc_autocorr_gamma(const double *values, int64_t n_samples, int n_temps, int max_lag, double *gamma_out) {
    autocorr *a = ac_new(max_lag, n_temps);
    for (int64_t i = 0; i < n_samples; i++) ac_push(a, values + (size_t)i * n_temps);
    ac_finish(a, gamma_out);
    ac_free(a);
}

int orc_equil_checkpoints(int64_t n_sweeps, int64_t *out) { /* equilibration.rs:18-29 */
    int n = 0;
    int64_t last = -1;
    for (int64_t p = 128; p < n_sweeps; p *= 2) { if (out) out[n] = p; last = p; n++; }
    if (last != n_sweeps) { if (out) out[n] = n_sweeps; n++; }
    return n;
}

typedef struct {
    double *mags, *mags2, *mags4, *energies, *energies2; /* averages [T] */
    double *ov[6];                                        /* averages [T] */
    uint64_t *hist;                                       /* [T][N+1] */
    double *ql, *ql2;                                     /* [T][N+1] */
    double *m2_tau, *q2_tau;                              /* [T] each when autocorrelation is on, else NULL */
    double *equil;                                        /* [n_ckpt][2][T] when the equilibration diagnostic is on */
} real_result;

/* mcmc/tempering.rs:73-102 with the draw supplied by the caller */
static int attempt_edge(const float *energies, const float *temperatures, int64_t *system_ids,
                        int64_t n_spins, float log_rand, int temp_id, int64_t *left, int64_t *right) {
    float temp_1 = temperatures[temp_id];
    float temp_2 = temperatures[temp_id + 1];
    float energy_1 = energies[system_ids[temp_id]];
    float energy_2 = energies[system_ids[temp_id + 1]];
    *left = system_ids[temp_id];
    *right = system_ids[temp_id + 1];
    float delta = (float)n_spins * (energy_2 - energy_1) * (1.0f / temp_1 - 1.0f / temp_2);
    int accepted = delta >= log_rand;
    if (accepted) {
        int64_t t = system_ids[temp_id];
        system_ids[temp_id] = system_ids[temp_id + 1];
        system_ids[temp_id + 1] = t;
    }
    return accepted;
}

static void run_realization(orc_sim *sim, int64_t ridx, const orc_config *cfg, real_result *res,
                            uint32_t sweep0, uint32_t pt_event0) {
    realization *re = &sim->reals[ridx];
    const orc_lattice *lat = sim->lat;
    int T = sim->n_temps, R = sim->n_replicas, S = T * R, z = lat->n_neighbors;
    int64_t N = lat->n_spins;
    int n_pairs = R / 2;
    int64_t bins = N + 1;
    int philox = sim->rng_mode != ORC_RNG_XOSHIRO;
    int msc = sim->rng_mode == ORC_RNG_PHILOX_MSC;
    /* MSC layout: the 32 samples of a word share one key (RNG-SPEC) */
    uint64_t sweep_key = msc ? orc_splitmix64(sim->cur_seed ^ orc_splitmix64(ORC_MSC_KEY_DOMAIN ^ (uint64_t)((sim->sample_offset + ridx) >> 5)))
                             : re->base_seed;

    /* simulation/mod.rs:190-198: lookup only for Metropolis (reference); the Philox modes also use the
     * integer heat-bath table for +-J Gibbs (identical decisions to the log form on the 24-bit grid). */
    uint32_t *table = NULL;
    if (!cfg->force_log_form && (cfg->sweep_mode == ORC_SWEEP_METROPOLIS || philox))
        table = build_table(re->couplings, N * z, re->temperatures, S, z, cfg->sweep_mode);

    double *s_m = calloc((size_t)T * 5, sizeof(double));
    double *s_m2 = s_m + T, *s_m4 = s_m + 2 * T, *s_e = s_m + 3 * T, *s_e2 = s_m + 4 * T;
    int64_t stat_count = 0;
    double *s_ov = calloc((size_t)T * 6 + 1, sizeof(double));
    int64_t ov_count = 0;
    int64_t *msums = calloc((size_t)S + 1, sizeof(int64_t));
    uint32_t pt_event = pt_event0;
    /* mod.rs:373-383, equilibration.rs: running sums of the replica-mean energy and pair-mean link overlap of EVERY sweep */
    int equil = cfg->equil_diag != 0;
    int64_t ckpts[80];
    int n_ckpt = equil ? orc_equil_checkpoints(cfg->n_sweeps, ckpts) : 0, next_ckpt = 0;
    int64_t eq_count = 0;
    double *eq_sum = equil ? calloc((size_t)T * 2, sizeof(double)) : NULL; /* energy row, link-overlap row */
    float *diag_e = equil ? calloc((size_t)T * 2, sizeof(float)) : NULL, *diag_ql = equil ? diag_e + T : NULL;
    /* mod.rs:341-371: autocorrelation accumulators over the recorded sweeps */
    autocorr *m2_acc = NULL, *q2_acc = NULL;
    double *m2_ac_buf = NULL, *q2_ac_buf = NULL;
    if (cfg->autocorr_max_lag > 0) {
        int64_t n_meas = cfg->n_sweeps > cfg->warmup_sweeps ? cfg->n_sweeps - cfg->warmup_sweeps : 0;
        int64_t k = cfg->autocorr_max_lag < n_meas / 4 ? cfg->autocorr_max_lag : n_meas / 4;
        if (k < 1) k = 1;
        m2_acc = ac_new((int)k, T);
        m2_ac_buf = calloc((size_t)T, sizeof(double));
        if (n_pairs > 0) { q2_acc = ac_new((int)k, T); q2_ac_buf = calloc((size_t)T, sizeof(double)); }
    }

    for (int64_t sweep_id = 0; sweep_id < cfg->n_sweeps; sweep_id++) {
        int record = sweep_id >= cfg->warmup_sweeps; /* mod.rs:410 */
        uint32_t sweep_index = sweep0 + (uint32_t)sweep_id;

        /* mod.rs:412-432 */
        if (!philox) {
            int z_ = z, offset = 2 * z_, width = 4 * z_ + 1;
            for (int slot = 0; slot < S; slot++) {
                int64_t sys = re->system_ids[slot];
                int8_t *s = re->spins + sys * N;
                uint64_t *rng = re->rngs + sys * 4;
                const uint32_t *row = table ? table + (size_t)slot * width : NULL;
                for (int64_t i = 0; i < N; i++) {
                    float h = local_field(lat, s, re->couplings, i);
                    attempt(s, i, h, xo_draw24(rng), re->temperatures[slot], cfg->sweep_mode, row, offset);
                }
            }
        } else {
            sweep_philox_impl(lat, re->spins, re->couplings, re->temperatures, re->system_ids, S, sim->colour,
                              sim->order, sim->rank, sweep_key, sweep_index, cfg->sweep_mode, table, msc,
                              sim->rng_mode == ORC_RNG_PHILOX_PACKED ? 1 : sim->rng_mode == ORC_RNG_PHILOX_SYSQ ? 2 : 0);
        }

        /* mod.rs:434-470: FK cluster update of every slot's system, after the sweep, before the measurements */
        if (cfg->cluster_interval > 0 && sweep_id % cfg->cluster_interval == 0)
            for (int slot = 0; slot < S; slot++) {
                int64_t sys = re->system_ids[slot];
                orc_fk_update(lat, re->spins + sys * N, re->couplings, re->temperatures[slot], re->base_seed, sweep_index,
                              (uint32_t)sys, cfg->cluster_wolff);
            }

        int pt_this_sweep = cfg->pt_interval > 0 && sweep_id % cfg->pt_interval == 0; /* mod.rs:486-488 */

        /* mod.rs:492-509 */
        if (record || pt_this_sweep || equil) orc_energies_mags(lat, re->spins, re->couplings, S, re->energies, record ? msums : NULL);

        if (equil) { /* mod.rs:511-525 */
            for (int t = 0; t < T; t++) { diag_e[t] = 0.0f; diag_ql[t] = 0.0f; }
            for (int r = 0; r < R; r++)
                for (int t = 0; t < T; t++) diag_e[t] += re->energies[re->system_ids[r * T + t]];
            float inv = 1.0f / (float)R;
            for (int t = 0; t < T; t++) diag_e[t] *= inv;
        }

        /* mod.rs:527-529 -> statistics/overlap.rs:251-333 (system_ids BEFORE this sweep's PT) */
        if ((record || equil) && n_pairs > 0) { /* mod.rs:527-529 */
            if (q2_ac_buf && record) for (int t = 0; t < T; t++) q2_ac_buf[t] = 0.0; /* overlap.rs:255-257 */
            for (int p = 0; p < n_pairs; p++) {
                for (int t = 0; t < T; t++) {
                    int64_t sa = re->system_ids[(2 * p) * T + t];
                    int64_t sb = re->system_ids[(2 * p + 1) * T + t];
                    int64_t dot_spin, dot_link;
                    orc_overlap_dots(lat, re->spins + sa * N, re->spins + sb * N, &dot_spin, &dot_link);
                    float ql = (float)dot_link / (float)(N * z);
                    if (equil) diag_ql[t] += ql; /* overlap.rs:285-287 */
                    if (!record) continue;       /* overlap.rs:288-290 */
                    float q = (float)dot_spin / (float)N;
                    float q2 = q * q;
                    float ql2 = ql * ql;
                    s_ov[0 * T + t] += (double)q;
                    s_ov[1 * T + t] += (double)q2;
                    s_ov[2 * T + t] += (double)(q2 * q2);
                    s_ov[3 * T + t] += (double)ql;
                    s_ov[4 * T + t] += (double)ql2;
                    s_ov[5 * T + t] += (double)(ql2 * ql2);
                    int64_t idx = (dot_spin + N) / 2;
                    res->hist[t * bins + idx] += 1;
                    res->ql[t * bins + idx] += (double)ql;
                    res->ql2[t * bins + idx] += (double)(ql * ql);
                    if (q2_ac_buf) q2_ac_buf[t] += (double)q2; /* overlap.rs:314-316 */
                }
                if (record) ov_count++;
            }
            if (equil) { /* overlap.rs:327-332 */
                float inv = 1.0f / (float)n_pairs;
                for (int t = 0; t < T; t++) diag_ql[t] *= inv;
            }
        }
        if (equil) { /* mod.rs:531-541 -> equilibration.rs:43-58 */
            eq_count++;
            for (int t = 0; t < T; t++) { eq_sum[t] += (double)diag_e[t]; eq_sum[T + t] += (double)diag_ql[t]; }
            if (next_ckpt < n_ckpt && eq_count == ckpts[next_ckpt]) {
                double c = (double)eq_count;
                for (int t = 0; t < 2 * T; t++) res->equil[(size_t)next_ckpt * 2 * T + t] = eq_sum[t] / c;
                next_ckpt++;
            }
        }

        /* mod.rs:543-578 + statistics/stats.rs:17-27 */
        if (record) {
            if (m2_ac_buf) for (int t = 0; t < T; t++) m2_ac_buf[t] = 0.0; /* mod.rs:551-553 */
            for (int r = 0; r < R; r++) {
                for (int t = 0; t < T; t++) {
                    int64_t sys = re->system_ids[r * T + t];
                    float mag = (float)msums[sys] / (float)N;
                    float m2 = mag * mag;
                    float m4 = m2 * m2;
                    float e = re->energies[sys];
                    if (m2_ac_buf) m2_ac_buf[t] += (double)m2; /* mod.rs:568-570 */
                    s_m[t] += (double)mag;
                    s_m2[t] += (double)m2;
                    s_m4[t] += (double)m4;
                    s_e[t] += (double)e;
                    s_e2[t] += (double)e * (double)e; /* powi(2) in f64 */
                }
                stat_count++;
            }
            if (m2_acc) { /* mod.rs:580-586 */
                double inv = 1.0 / (double)R;
                for (int t = 0; t < T; t++) m2_ac_buf[t] *= inv;
                ac_push(m2_acc, m2_ac_buf);
            }
            if (q2_acc) { /* mod.rs:588-594 */
                double inv = 1.0 / (double)n_pairs;
                for (int t = 0; t < T; t++) q2_ac_buf[t] *= inv;
                ac_push(q2_acc, q2_ac_buf);
            }
        }

        /* mod.rs:596-746: overlap cluster move after the measurements; energies are refreshed only for PT (:748-756) */
        if (cfg->overlap_cluster_interval > 0 && sweep_id % cfg->overlap_cluster_interval == 0) {
            for (int t = 0; t < T; t++)
                orc_houdayer_slot(lat, re->spins, re->system_ids, T, R, t, msc ? sweep_key : re->base_seed, sweep_index,
                                  cfg->overlap_cluster_wolff); /* multispin layout: lane-uniform draws use the group key */
            if (pt_this_sweep) orc_energies_mags(lat, re->spins, re->couplings, S, re->energies, NULL);
        }

        /* mod.rs:748-796 */
        if (pt_this_sweep) {
            int first_parity = re->pt.next_parity;
            for (int r = 0; r < R; r++) {
                int off = r * T;
                int64_t *sid = re->system_ids + off;
                const float *tsl = re->temperatures + off;
                if (T < 2) continue;
                if (cfg->pt_schedule == ORC_PT_SINGLE_RANDOM_EDGE) {
                    int edge; float log_rand;
                    if (!philox) { /* tempering.rs:33, :89 with rngs[offset] (mod.rs:779) */
                        uint64_t *rng = re->rngs + (size_t)off * 4;
                        edge = (int)xo_gen_range_usize(rng, 0, (uint64_t)(T - 1));
                        log_rand = logf(u24_to_f32(xo_draw24(rng)));
                    } else { /* RNG-SPEC PT domain */
                        uint32_t ctr[4] = {0xFFFFFFFFu, pt_event, (uint32_t)r, ORC_TAG_PT};
                        uint32_t k[2] = {(uint32_t)re->base_seed, (uint32_t)(re->base_seed >> 32)}, o[4];
                        orc_philox(ctr, k, o);
                        edge = (int)(((uint64_t)o[1] * (uint64_t)(T - 1)) >> 32);
                        log_rand = logf(u24_to_f32(o[0] >> 8));
                    }
                    int64_t left, right;
                    int acc = attempt_edge(re->energies, tsl, sid, N, log_rand, edge, &left, &right);
                    pt_record_attempt(&re->pt, edge, acc, left, right);
                } else {
                    int parities[2] = {first_parity, 1 - first_parity};
                    for (int pi = 0; pi < 2; pi++)
                        for (int edge = parities[pi]; edge < T - 1; edge += 2) {
                            float log_rand;
                            if (!philox) {
                                log_rand = logf(u24_to_f32(xo_draw24(re->rngs + (size_t)off * 4)));
                            } else {
                                uint32_t ctr[4] = {(uint32_t)edge, pt_event, (uint32_t)r, ORC_TAG_PT};
                                uint32_t k[2] = {(uint32_t)re->base_seed, (uint32_t)(re->base_seed >> 32)}, o[4];
                                orc_philox(ctr, k, o);
                                log_rand = logf(u24_to_f32(o[0] >> 8));
                            }
                            int64_t left, right;
                            int acc = attempt_edge(re->energies, tsl, sid, N, log_rand, edge, &left, &right);
                            pt_record_attempt(&re->pt, edge, acc, left, right);
                        }
                }
            }
            if (cfg->pt_schedule == ORC_PT_FULL_LADDER) re->pt.next_parity = 1 - re->pt.next_parity; /* mod.rs:793-795 */
            pt_event++;
        }
    }

    /* statistics/stats.rs:29-35: average = aggregate / count (aggregate itself if count == 0) */
    double c = stat_count > 0 ? (double)stat_count : 1.0;
    for (int t = 0; t < T; t++) {
        res->mags[t] = s_m[t] / c; res->mags2[t] = s_m2[t] / c; res->mags4[t] = s_m4[t] / c;
        res->energies[t] = s_e[t] / c; res->energies2[t] = s_e2[t] / c;
    }
    double oc = ov_count > 0 ? (double)ov_count : 1.0;
    for (int k = 0; k < 6; k++)
        for (int t = 0; t < T; t++) res->ov[k][t] = s_ov[k * T + t] / oc;
    /* mod.rs:825-832: tau[t] = sokal_tau(gamma[t]) */
    if (m2_acc) {
        int L1 = m2_acc->max_lag + 1;
        double *gamma = malloc(sizeof(double) * (size_t)T * L1);
        ac_finish(m2_acc, gamma);
        for (int t = 0; t < T; t++) res->m2_tau[t] = orc_sokal_tau(gamma + (size_t)t * L1, L1);
        if (q2_acc) {
            ac_finish(q2_acc, gamma);
            for (int t = 0; t < T; t++) res->q2_tau[t] = orc_sokal_tau(gamma + (size_t)t * L1, L1);
        }
        free(gamma);
    }
    ac_free(m2_acc); ac_free(q2_acc); free(m2_ac_buf); free(q2_ac_buf);
    free(eq_sum); free(diag_e);
    free(s_m); free(s_ov); free(msums); free(table);
}

typedef struct {
    orc_sim *sim;
    const orc_config *cfg;
    real_result *rr;
    uint32_t sweep0, pt0;
    int64_t D;
    int64_t next; /* dynamic schedule: one realization per grab */
} worker_ctx;

static void *worker_main(void *arg) {
    worker_ctx *ctx = arg;
    for (;;) {
        int64_t d = __atomic_fetch_add(&ctx->next, 1, __ATOMIC_RELAXED);
        if (d >= ctx->D) break;
        run_realization(ctx->sim, d, ctx->cfg, &ctx->rr[d], ctx->sweep0, ctx->pt0);
    }
    return NULL;
}

/* config.rs:180-247 (hot-path subset) */
static int validate(const orc_config *cfg) {
    if (cfg->n_sweeps < 1) { set_err("n_sweeps must be >= 1"); return -1; }
    if (cfg->warmup_sweeps > cfg->n_sweeps) { set_err("warmup_sweeps must be <= n_sweeps"); return -1; }
    if (cfg->pt_interval < 0) { set_err("pt_interval must be >= 1"); return -1; }
    if (cfg->cluster_interval < 0) { set_err("cluster_update_interval must be >= 1"); return -1; }
    return 0;
}

/* simulation/mod.rs:865-939 + statistics/results.rs:165-180, 250-259 + statistics/overlap.rs:106-152 */
int orc_sim_sample(orc_sim *sim, const orc_config *cfg, orc_results *out) {
    if (validate(cfg) != 0) return -1;
    if (cfg->overlap_cluster_interval > 0 && sim->n_replicas < 2) {
        set_err("overlap cluster requires n_replicas >= max group_size"); /* mod.rs:207-213 */
        return -1;
    }
    if ((cfg->cluster_interval > 0 && sim->rng_mode != ORC_RNG_PHILOX && sim->rng_mode != ORC_RNG_PHILOX_PACKED &&
         sim->rng_mode != ORC_RNG_PHILOX_SYSQ) ||
        (cfg->overlap_cluster_interval > 0 && sim->rng_mode == ORC_RNG_XOSHIRO)) {
        set_err("cluster updates are restated for the RNG-SPEC modes only (FK: int8 mode)");
        return -1;
    }
    int T = sim->n_temps, R = sim->n_replicas;
    int64_t N = sim->lat->n_spins, bins = N + 1, D = sim->n_real;
    int n_pairs = R / 2;
    real_result *rr = calloc((size_t)D, sizeof(real_result));
    for (int64_t d = 0; d < D; d++) {
        rr[d].mags = calloc((size_t)T * 11 + 1, sizeof(double));
        rr[d].mags2 = rr[d].mags + T; rr[d].mags4 = rr[d].mags + 2 * T;
        rr[d].energies = rr[d].mags + 3 * T; rr[d].energies2 = rr[d].mags + 4 * T;
        for (int k = 0; k < 6; k++) rr[d].ov[k] = rr[d].mags + (5 + k) * T;
        if (cfg->autocorr_max_lag > 0) {
            rr[d].m2_tau = calloc((size_t)T * 2, sizeof(double));
            rr[d].q2_tau = rr[d].m2_tau + T;
        }
        if (cfg->equil_diag) rr[d].equil = calloc((size_t)orc_equil_checkpoints(cfg->n_sweeps, NULL) * 2 * T, sizeof(double));
        if (n_pairs > 0) {
            if (out->ps_hist) {
                rr[d].hist = out->ps_hist + (size_t)d * T * bins;
                rr[d].ql = out->ps_ql_at_q_sum + (size_t)d * T * bins;
                rr[d].ql2 = out->ps_ql2_at_q_sum + (size_t)d * T * bins;
                memset(rr[d].hist, 0, sizeof(uint64_t) * (size_t)T * bins);
                memset(rr[d].ql, 0, sizeof(double) * (size_t)T * bins);
                memset(rr[d].ql2, 0, sizeof(double) * (size_t)T * bins);
            } else {
                rr[d].hist = calloc((size_t)T * bins, sizeof(uint64_t));
                rr[d].ql = calloc((size_t)T * bins, sizeof(double));
                rr[d].ql2 = calloc((size_t)T * bins, sizeof(double));
            }
        }
    }
    uint32_t sweep0 = sim->sweep_counter, pt0 = sim->pt_event_counter;
    /* mod.rs:887-903: threads over realizations (rayon par_iter_mut there; pthreads here) */
    int nthreads = cfg->n_threads > 1 ? cfg->n_threads : 1;
    if (nthreads > D) nthreads = (int)D;
    worker_ctx ctx = {sim, cfg, rr, sweep0, pt0, D, 0};
    if (nthreads <= 1) {
        worker_main(&ctx);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker_main, &ctx);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
        free(th);
    }

    sim->sweep_counter += (uint32_t)cfg->n_sweeps;
    if (cfg->pt_interval > 0) sim->pt_event_counter += (uint32_t)((cfg->n_sweeps + cfg->pt_interval - 1) / cfg->pt_interval);

    /* results.rs:165-180 then :250-259: sum over realizations in order, divide by D */
    double n = (double)D;
    double *dst[5] = {out->mags, out->mags2, out->mags4, out->energies, out->energies2};
    for (int k = 0; k < 5; k++) {
        for (int t = 0; t < T; t++) dst[k][t] = 0.0;
        for (int64_t d = 0; d < D; d++)
            for (int t = 0; t < T; t++) dst[k][t] += rr[d].mags[k * T + t];
        for (int t = 0; t < T; t++) dst[k][t] /= n;
    }
    if (n_pairs > 0) {
        double *od[6] = {out->overlap, out->overlap2, out->overlap4, out->link_overlap, out->link_overlap2,
                         out->link_overlap4};
        for (int k = 0; k < 6; k++) {
            for (int t = 0; t < T; t++) od[k][t] = 0.0;
            for (int64_t d = 0; d < D; d++)
                for (int t = 0; t < T; t++) od[k][t] += rr[d].ov[k][t];
            for (int t = 0; t < T; t++) od[k][t] /= n;
        }
        memset(out->hist, 0, sizeof(uint64_t) * (size_t)T * bins);
        memset(out->ql_at_q_sum, 0, sizeof(double) * (size_t)T * bins);
        memset(out->ql2_at_q_sum, 0, sizeof(double) * (size_t)T * bins);
        for (int64_t d = 0; d < D; d++)
            for (int64_t i = 0; i < (int64_t)T * bins; i++) {
                out->hist[i] += rr[d].hist[i];
                out->ql_at_q_sum[i] += rr[d].ql[i];
                out->ql2_at_q_sum[i] += rr[d].ql2[i];
            }
    }
    if (out->edge_attempts) {
        int n_edges = T > 0 ? T - 1 : 0;
        for (int64_t d = 0; d < D; d++) {
            memcpy(out->edge_attempts + d * n_edges, sim->reals[d].pt.edge_attempts, sizeof(uint64_t) * (size_t)n_edges);
            memcpy(out->edge_acceptances + d * n_edges, sim->reals[d].pt.edge_acceptances, sizeof(uint64_t) * (size_t)n_edges);
            memcpy(out->round_trips + d * R * T, sim->reals[d].pt.round_trips, sizeof(uint64_t) * (size_t)R * T);
        }
    }
    if (cfg->autocorr_max_lag > 0) { /* results.rs:217-231, 269-274: sum over realizations in order, divide by D */
        double *td[2] = {out->mags2_tau, n_pairs > 0 ? out->overlap2_tau : NULL};
        for (int k = 0; k < 2; k++) {
            if (!td[k]) continue;
            for (int t = 0; t < T; t++) td[k][t] = 0.0;
            for (int64_t d = 0; d < D; d++)
                for (int t = 0; t < T; t++) td[k][t] += rr[d].m2_tau[k * T + t];
            for (int t = 0; t < T; t++) td[k][t] /= n;
        }
        if (out->ps_taus)
            for (int64_t d = 0; d < D; d++) memcpy(out->ps_taus + (size_t)d * 2 * T, rr[d].m2_tau, sizeof(double) * 2 * (size_t)T);
    }
    if (cfg->equil_diag) { /* results.rs:231-247, 275-282 */
        int n_ckpt = orc_equil_checkpoints(cfg->n_sweeps, NULL);
        double *ed[2] = {out->equil_energy_avg, out->equil_link_overlap_avg};
        for (int k = 0; k < 2; k++) {
            if (!ed[k]) continue;
            for (int c = 0; c < n_ckpt; c++)
                for (int t = 0; t < T; t++) {
                    double acc = 0.0;
                    for (int64_t d = 0; d < D; d++) acc += rr[d].equil[((size_t)c * 2 + k) * T + t];
                    ed[k][(size_t)c * T + t] = acc / n;
                }
        }
        if (out->ps_equil)
            for (int64_t d = 0; d < D; d++)
                memcpy(out->ps_equil + (size_t)d * n_ckpt * 2 * T, rr[d].equil, sizeof(double) * (size_t)n_ckpt * 2 * T);
    }
    if (out->ps_means) /* [D][11][T]: rr[d].mags is one block of 11 rows (mags..energies2, then the six overlap rows) */
        for (int64_t d = 0; d < D; d++) memcpy(out->ps_means + (size_t)d * 11 * T, rr[d].mags, sizeof(double) * 11 * (size_t)T);
    for (int64_t d = 0; d < D; d++) {
        free(rr[d].mags);
        free(rr[d].m2_tau);
        free(rr[d].equil);
        if (n_pairs > 0 && !out->ps_hist) { free(rr[d].hist); free(rr[d].ql); free(rr[d].ql2); }
    }
    free(rr);
    return 0;
}
