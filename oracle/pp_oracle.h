/*
 * pp_oracle.h — CPU oracle for the peapods spin-sim sweep path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (PeaBrane/peapods v0.2.1, Rust) single-spin-flip sweep engine.  It is the
 * checker the CUDA path is compared against and the CPU baseline bench.py
 * times; nothing under peapods_b200/ may link, import or call it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 *
 * Parity status
 *   pinned   : lattice tables, energy/magnetisation, LUT cut-offs, PT edge order,
 *              round-trip counter, seeding replay — against the reference's own
 *              KATs (geometry/lattice.rs:116-184, spins/energy.rs:117-147,
 *              mcmc/sweep.rs:346-380, mcmc/tempering.rs:110-138,
 *              simulation/realization.rs:267-302).
 *   unpinned : the raw xoshiro256** / rand-0.8.5 stream ("parity unpinned":
 *              rand 0.8.5, rand_core 0.6.4, rand_xoshiro 0.6.0 are not vendored
 *              under /root/reference and no reference test records a raw draw
 *              or a golden spin configuration).  xoshiro256** itself is checked
 *              against its authors' published vector; rand's adaptor semantics
 *              are restated from the crate's documented algorithm.
 *
 * Two visit-order / RNG modes:
 *   ORC_RNG_XOSHIRO     reference-faithful: typewriter order, one xoshiro256**
 *                       stream per system (mcmc/sweep.rs:51-97, parallel.rs:27-33).
 *   ORC_RNG_PHILOX      the build's deterministic contract (RNG-SPEC):
 *                       colour classes ascending, counter-based Philox4x32-10
 *                       keyed by (seed, sweep, site-rank, stream); per-site
 *                       arithmetic and acceptance rule are the reference's.
 *   ORC_RNG_PHILOX_PACKED as PHILOX with the sweep draws in the packed mapping (32 ranks per six calls: orc_draw24_packed)
 *   ORC_RNG_PHILOX_SYSQ as PHILOX with the sweep draws in the system-quad mapping (kernels that keep the same site of 32 systems of a
 *                       realization in one word): counter = {colour rank, sweep index, system >> 2, ORC_TAG_SWEEP_SYSQ | colour},
 *                       the draw of system s is out[s & 3] >> 8
 *   ORC_RNG_PHILOX_MSC  as PHILOX but one draw shared by the 32 disorder samples
 *                       of a multispin word (stream = replica*T + slot).
 */
#ifndef PP_ORACLE_H
#define PP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_DIMS 8

enum { ORC_RNG_XOSHIRO = 0, ORC_RNG_PHILOX = 1, ORC_RNG_PHILOX_MSC = 2, ORC_RNG_PHILOX_PACKED = 3, ORC_RNG_PHILOX_SYSQ = 4 };
enum { ORC_SWEEP_METROPOLIS = 0, ORC_SWEEP_GIBBS = 1 };
enum { ORC_PT_SINGLE_RANDOM_EDGE = 0, ORC_PT_FULL_LADDER = 1 };

/* RNG-SPEC domain tags (counter word 3; low 16 bits carry the colour). */
#define ORC_TAG_INIT      0x00010000u
#define ORC_TAG_SWEEP     0x00020000u
#define ORC_TAG_PT        0x00030000u
#define ORC_TAG_SWEEP_MSC 0x00040000u
#define ORC_TAG_FK_BOND   0x00050000u  /* counter = {bond >> 2, sweep index, system id, tag}; bond = site * z' + direction */
#define ORC_TAG_FK_FLIP   0x00060000u  /* counter = {root >> 2 | 0xFFFFFFFF (Wolff seed), sweep index, system id, tag} */
#define ORC_TAG_OC_PAIR   0x00070000u  /* replica shuffle at a temperature: counter = {step, sweep index, slot t, tag} */
#define ORC_TAG_OC_SEED   0x00080000u  /* Wolff seed scores: counter = {site >> 2, sweep index, t * P + g, tag} */
#define ORC_TAG_OC_FLIP   0x00090000u  /* cluster coins:     counter = {root >> 2, sweep index, t * P + g, tag} */
#define ORC_TAG_SWEEP_SYSQ 0x000B0000u  /* counter = {colour rank, sweep index, system id >> 2, tag | colour}; system s draws out[s & 3] >> 8 */
#define ORC_TAG_SWEEP_PACKED 0x000A0000u  /* counter = {rank >> 5, sweep index, system id, tag | call << 8 | colour}, call = 0..5 */
#define ORC_MSC_KEY_DOMAIN 0x6D73635F67726F75ull

typedef struct orc_lattice orc_lattice;
typedef struct orc_sim orc_sim;

typedef struct {
    int64_t n_sweeps;
    int64_t warmup_sweeps;
    int32_t sweep_mode;   /* ORC_SWEEP_* */
    int64_t pt_interval;  /* 0 = no parallel tempering */
    int32_t pt_schedule;  /* ORC_PT_* */
    int32_t n_threads;    /* threads over realizations (<=1: sequential) */
    int32_t force_log_form; /* 1: never use the +-J lookup (test of LUT == log form) */
    int64_t autocorr_max_lag; /* 0 = off; simulation/mod.rs:342-344: clamped to [1, recorded sweeps / 4] */
    int64_t cluster_interval; /* 0 = off; Fortuin-Kasteleyn cluster update every this many sweeps (simulation/mod.rs:434-470) */
    int32_t cluster_wolff;    /* 0: Swendsen-Wang (every cluster flips with probability 1/2); 1: Wolff (the seed's cluster flips) */
    int64_t overlap_cluster_interval; /* 0 = off; Houdayer isoenergetic cluster move (group size 2) every this many sweeps */
    int32_t overlap_cluster_wolff;    /* 1: flip the cluster of one drawn active site (default); 0: every multi-site cluster w.p. 1/2 */
    int32_t equil_diag;       /* 1: equilibration diagnostic (statistics/equilibration.rs; energies + link overlaps every sweep) */
} orc_config;

typedef struct {
    /* f64[T] each */
    double *mags, *mags2, *mags4, *energies, *energies2;
    /* f64[T] each; only written when n_replicas >= 2 */
    double *overlap, *overlap2, *overlap4, *link_overlap, *link_overlap2, *link_overlap4;
    /* [T][N+1]; summed over realizations */
    uint64_t *hist;
    double *ql_at_q_sum, *ql2_at_q_sum;
    /* [D][T][N+1]; may be NULL */
    uint64_t *ps_hist;
    double *ps_ql_at_q_sum, *ps_ql2_at_q_sum;
    /* [D][T-1], [D][T-1], [D][R][T]; may be NULL */
    uint64_t *edge_attempts, *edge_acceptances, *round_trips;
    /* [D][11][T] per-realization averages (mags, mags2, mags4, energies, energies2, q, q2, q4, ql, ql2, ql4); may be NULL */
    double *ps_means;
    /* [T] each: integrated autocorrelation times of m^2 and q^2, mean over realizations (statistics/results.rs:217-272);
     * written when autocorr_max_lag > 0 (overlap2_tau only when n_replicas >= 2); may be NULL */
    double *mags2_tau, *overlap2_tau;
    /* [D][2][T] per-realization taus (m^2 row, q^2 row); may be NULL */
    double *ps_taus;
    /* equil_diag: running averages at the checkpoints of orc_equil_checkpoints(n_sweeps): [n_ckpt][T] each, mean over
     * realizations (statistics/results.rs:231-247, 275-282); may be NULL */
    double *equil_energy_avg, *equil_link_overlap_avg;
    /* [D][n_ckpt][2][T] per-realization checkpoints (energy row, link-overlap row); may be NULL */
    double *ps_equil;
} orc_results;

/* ---- RNG primitives ---------------------------------------------------- */
uint64_t orc_splitmix64(uint64_t v);                       /* realization.rs:9-15 */
uint64_t orc_child_seed(uint64_t root, uint64_t domain, uint64_t index); /* realization.rs:17-19 */
uint64_t orc_realization_seed(uint64_t root, uint64_t r);  /* src/lib.rs:30-32 */
void orc_xoshiro_seed_from_u64(uint64_t s[4], uint64_t seed);
uint64_t orc_xoshiro_next_u64(uint64_t s[4]);
#define ORC_PHILOX_ROUNDS 7  /* RNG-SPEC v2 */
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4]);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);  /* published test vectors */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);         /* the RNG-SPEC generator (R = 7) */
uint32_t orc_draw24(uint64_t key, uint32_t c0_index, uint32_t c1, uint32_t c2, uint32_t c3);
uint32_t orc_draw24_packed(uint64_t key, uint32_t rank, uint32_t sweep, uint32_t stream, uint32_t tag_colour);

/* ---- lattice (geometry/lattice.rs:44-109) ------------------------------ */
orc_lattice *orc_lattice_new(int n_dims, const int64_t *shape, int n_offsets,
                             const int64_t *offsets /* NULL: hypercubic */);
void orc_lattice_free(orc_lattice *lat);
int64_t orc_lattice_n_spins(const orc_lattice *lat);
int orc_lattice_n_neighbors(const orc_lattice *lat);
int64_t orc_lattice_stride(const orc_lattice *lat, int d);
uint32_t orc_neighbor_fwd(const orc_lattice *lat, int64_t i, int d);
uint32_t orc_neighbor_bwd(const orc_lattice *lat, int64_t i, int d);
/* 1 iff no site shares a colour with any of its 2z' neighbours */
int orc_colouring_is_valid(const orc_lattice *lat, const uint16_t *colour);

/* ---- acceptance tables (mcmc/sweep.rs:99-167) -------------------------- */
/* returns 0 and fills table[n_temps*(4z'+1)] iff eligible, else -1 (fail closed) */
int orc_metropolis_lookup(const float *couplings, int64_t n_couplings, const float *temps,
                          int n_temps, int n_neighbors, uint32_t *table);
uint32_t orc_metropolis_accepted_count(float temperature, int32_t energy_change);
int orc_metropolis_legacy_accepts(float temperature, int32_t energy_change, uint32_t draw);
/* heat-bath analogue for integer fields: #{draw : ec >= (T/2) ln(u/(1-u))} */
uint32_t orc_gibbs_accepted_count(float temperature, int32_t energy_change);
int orc_gibbs_legacy_accepts(float temperature, int32_t energy_change, uint32_t draw);

/* ---- operator-level restatements --------------------------------------- */
/* spins/energy.rs:78-110: e[sys] = (sum_i sum_d s_i s_fwd J)/N (f32 sequential), M = sum s_i */
void orc_energies_mags(const orc_lattice *lat, const int8_t *spins, const float *couplings,
                       int64_t n_systems, float *energies, int64_t *mags /* may be NULL */);
/* statistics/overlap.rs:259-281: integer dots for one (a, b) configuration pair */
void orc_overlap_dots(const orc_lattice *lat, const int8_t *spins_a, const int8_t *spins_b,
                      int64_t *dot_spin, int64_t *dot_link);
/* mcmc/sweep.rs:220-284 with xoshiro streams in typewriter order. rng_states: u64[S][4] */
void orc_sweep_xoshiro(const orc_lattice *lat, int8_t *spins, const float *couplings,
                       const float *temperatures, const int64_t *system_ids, int64_t n_systems,
                       uint64_t *rng_states, int sweep_mode, int use_lookup);
/* same per-site arithmetic, colour order + Philox draws (RNG-SPEC).
 * stream_is_slot = 0: stream id = system id (int8 layout); 1: stream id = slot (MSC layout) */
void orc_sweep_philox(const orc_lattice *lat, int8_t *spins, const float *couplings,
                      const float *temperatures, const int64_t *system_ids, int64_t n_systems,
                      const uint16_t *colour, uint64_t key, uint32_t sweep_index,
                      int sweep_mode, int use_lookup, int stream_is_slot);
/* mcmc/tempering.rs:45-70: order in which full-ladder attempts edges */
int orc_full_ladder_edges(int n_temps, int first_parity, int32_t *edges_out);

/* simulation/realization.rs:73-120 replayed over a caller-supplied attempt list */
void orc_pt_replay(int n_replicas, int n_temps, const float *temps, int n_attempts, const int32_t *edges,
                   const int32_t *accepted, const int64_t *left, const int64_t *right,
                   uint64_t *edge_attempts, uint64_t *edge_acceptances, uint64_t *round_trips);

/* ---- autocorrelation (statistics/autocorrelation.rs) -------------------- */
/* AutocorrAccum ring backend (:24-124, :166-199): push n_samples rows values[n_samples][n_temps], then finish():
 * gamma_out[n_temps][max_lag + 1] */
void orc_autocorr_gamma(const double *values, int64_t n_samples, int n_temps, int max_lag, double *gamma_out);
/* sokal_tau (:201-210) over gamma[0..n) */
double orc_sokal_tau(const double *gamma, int n);

/* ---- Fortuin-Kasteleyn cluster update (clusters/fk.rs:28-171, union-find path) under RNG-SPEC draws ---- */
/* #{draw in [0, 2^24) : draw * 2^-24 < 1 - expf(-2 / T)}: the bond probability of an aligned unit-coupling pair (fk.rs:108-114) */
uint32_t orc_fk_bond_count(float temperature);
/* one system: bonds between aligned neighbours with s_i s_j J > 0 are activated by their own draw, clusters are the connected
 * components (labelled by their smallest site), then SW flips each cluster whose root draw is < 1/2, Wolff flips the cluster of
 * the drawn seed site.  Unit couplings (|J| in {0, 1}) only. */
void orc_fk_update(const orc_lattice *lat, int8_t *spins, const float *couplings, float temperature, uint64_t key,
                   uint32_t sweep_index, uint32_t system_id, int wolff);

/* ---- Houdayer isoenergetic cluster move, group size 2 (clusters/overlap.rs:34-56, 146-339) under RNG-SPEC draws ----
 * One realization, temperature slot t: the R systems at the slot are shuffled (Fisher-Yates, draws of ORC_TAG_OC_PAIR) and paired;
 * for pair g the active sites are those where the two replicas differ, bonds join active neighbours, clusters are named by
 * their smallest site.  wolff: the active site with the smallest (score, index) is the seed (uniform over active sites) and its
 * cluster flips in both replicas; else every cluster of more than one site flips iff its root draw is < 1/2. */
void orc_houdayer_slot(const orc_lattice *lat, int8_t *spins /* [S][N] */, const int64_t *system_ids /* [S] */, int n_temps,
                       int n_replicas, int t, uint64_t key, uint32_t sweep_index, int wolff);

/* statistics/equilibration.rs:18-29: 128, 256, ... < n_sweeps, then n_sweeps.  Returns the count (out may be NULL). */
int orc_equil_checkpoints(int64_t n_sweeps, int64_t *out);

/* ---- full simulation (simulation/mod.rs:405-796, 865-939) -------------- */
orc_sim *orc_sim_new(int n_dims, const int64_t *shape, int n_offsets, const int64_t *offsets,
                     const float *couplings, int64_t n_realizations, const float *temps,
                     int n_temps, int n_replicas, uint64_t seed, int rng_mode,
                     const uint16_t *colour /* required for PHILOX modes */);
void orc_sim_free(orc_sim *sim);
void orc_sim_reset(orc_sim *sim, int has_seed, uint64_t seed); /* src/lib.rs:624-633 */
/* this sim holds realizations [offset, offset + D) of a larger run (seeds use global indices); re-initialises */
void orc_sim_set_sample_offset(orc_sim *sim, int64_t offset);
/* 0 ok; -1 invalid config (message via orc_last_error) */
int orc_sim_sample(orc_sim *sim, const orc_config *cfg, orc_results *out);
const int8_t *orc_sim_spins(const orc_sim *sim, int64_t realization);       /* [S*N] */
const int64_t *orc_sim_system_ids(const orc_sim *sim, int64_t realization); /* [S] */
const float *orc_sim_energies(const orc_sim *sim, int64_t realization);     /* [S] by system */
const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
