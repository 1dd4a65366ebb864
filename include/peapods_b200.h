/*
 * peapods_b200.h — C ABI of the B200-native spin-sim sweep engine.
 *
 * This is the drop-in boundary for the hot path of PeaBrane/peapods v0.2.1
 * (`spin-sim` single-spin-flip sweeps + parallel tempering + per-sweep energy /
 * magnetisation / overlap reductions).  Plain pointers and sizes only; every
 * buffer in a signature is HOST memory owned by the caller, the library copies
 * what it keeps.  No exceptions cross the boundary: every call returns a
 * pp_status and pp_last_error() gives the thread-local message.
 *
 * Reference interfaces replaced (paths under /root/reference/):
 *   pp_create        <- IsingSimulation::new          src/lib.rs:106-174
 *                       (Lattice::new/with_offsets    spin-sim/src/geometry/lattice.rs:31-93,
 *                        Realization::new             spin-sim/src/simulation/realization.rs:155-210)
 *   pp_sample        <- IsingSimulation::sample       src/lib.rs:176-333
 *                       -> run_sweep_parallel         spin-sim/src/simulation/mod.rs:865-939
 *                       -> run_sweep_loop_impl        spin-sim/src/simulation/mod.rs:405-796
 *   pp_get_spins     <- IsingSimulation::get_spins    src/lib.rs:620-622
 *   pp_reset         <- IsingSimulation::reset        src/lib.rs:624-633
 *   pp_op_sweep      <- metropolis_sweep / gibbs_sweep        spin-sim/src/mcmc/sweep.rs:220-284
 *   pp_op_energies_mags <- compute_energies_and_magnetizations_into  spin-sim/src/spins/energy.rs:59-76
 *   pp_op_overlap    <- OverlapAccum::collect (integer dots)  spin-sim/src/statistics/overlap.rs:259-281
 *   pp_op_pt         <- parallel_tempering(_full_ladder)      spin-sim/src/mcmc/tempering.rs:20-102
 *   pp_colouring     <- (no reference equivalent: the visit order of the checkerboard sweep)
 *   pp_nccl_unique_id, pp_model_desc.slab_* <- (no reference equivalent: one lattice across GPUs)
 *   pp_metropolis_lookup <- UnitCouplingMetropolisLookup::new spin-sim/src/mcmc/sweep.rs:102-159
 *
 * A handle is NOT thread-safe: one caller at a time (the reference takes &mut self).
 */
#ifndef PEAPODS_B200_H
#define PEAPODS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 4
#define PP_MAX_DIMS 8

typedef enum {
    PP_OK = 0,
    PP_ERR_INVALID = 1,     /* bad argument / config (reference: ValueError)            */
    PP_ERR_UNSUPPORTED = 2, /* valid in the reference, not on this path; nothing mutated */
    PP_ERR_INTERRUPTED = 3, /* interrupt flag seen (reference: KeyboardInterrupt)        */
    PP_ERR_CUDA = 4,
    PP_ERR_NCCL = 5,
    PP_ERR_OOM = 6
} pp_status;

enum { PP_SWEEP_METROPOLIS = 0, PP_SWEEP_GIBBS = 1 };              /* config.rs:3-20   */
enum { PP_PT_SINGLE_RANDOM_EDGE = 0, PP_PT_FULL_LADDER = 1 };      /* config.rs:61-79  */
enum { PP_LAYOUT_AUTO = 0, PP_LAYOUT_INT8 = 1, PP_LAYOUT_MSC = 2,
       PP_LAYOUT_SLAB = 3 /* one 3-D hypercubic ferromagnet, stride geometry (no tables), slab-decomposed along x0 */ };
#define PP_NCCL_ID_BYTES 128
enum { PP_COUPLINGS_ARRAY = 0, PP_COUPLINGS_FERRO = 1 };

typedef struct pp_sim pp_sim;

/* Constructor arguments (src/lib.rs:106-174). */
typedef struct {
    int32_t n_dims;
    const int64_t *shape;        /* [n_dims] */
    int32_t n_offsets;           /* 0: hypercubic (Lattice::new) */
    const int64_t *offsets;      /* [n_offsets][n_dims] forward neighbour offsets, or NULL */
    int32_t coupling_kind;       /* PP_COUPLINGS_ARRAY: `couplings` given; PP_COUPLINGS_FERRO: all +1, never materialised */
    const float *couplings;      /* [n_disorder][n_spins][n_neighbors], bond (i,d) owned by the lower site (lattice.rs:4-8) */
    int64_t n_disorder;          /* realizations held by THIS handle */
    int64_t sample_offset;       /* global index of this handle's first realization (multi-GPU shard; seeds use global indices) */
    const float *temperatures;   /* [n_temps] */
    int32_t n_temps;
    int32_t n_replicas;
    uint64_t seed;               /* dynamics seed (lib.rs:155; realization r uses splitmix64(seed ^ splitmix64(r)), lib.rs:30-32) */
    int32_t layout;              /* PP_LAYOUT_* ; AUTO = MSC when eligible (>=32 realizations, all couplings +-1) */
    int32_t device;              /* CUDA device ordinal */
    /* PP_LAYOUT_SLAB only: the lattice is cut along dimension 0 into slab_ranks slabs, one per process / GPU
     * (shape[0] must be a multiple of 2 * slab_ranks).  0 or 1 = the whole lattice on this device. */
    int32_t slab_ranks;
    int32_t slab_rank;           /* this process's slab, or -1: keep all slabs on this device (single-GPU emulation, tests) */
    const uint8_t *nccl_unique_id; /* [PP_NCCL_ID_BYTES] from pp_nccl_unique_id() on rank 0, broadcast by the host; NULL unless
                                      (slab_ranks > 1 and slab_rank >= 0) or system_ranks > 1 */
    /* int8 layout, ONE realization of many large systems (BASELINE configs[2]): the S = n_replicas * n_temps systems are split
     * over system_ranks processes / GPUs in contiguous blocks of S / system_ranks systems (the reference parallelises over
     * systems the same way: spin-sim/src/parallel.rs:36-40).  Every rank sweeps its own systems; energies and magnetisations are
     * all-gathered per measurement / exchange event, configurations per recorded sweep when n_replicas >= 2, and every rank
     * replays the same exchange decisions and statistics.  0 or 1 = off. */
    int32_t system_ranks;
    int32_t system_rank;
} pp_model_desc;

enum { PP_CLUSTER_SW = 0, PP_CLUSTER_WOLFF = 1 };  /* config.rs ClusterMode */

/* sample() arguments (src/lib.rs:176-284). Zero means "None" for the optional intervals. */
typedef struct {
    int64_t n_sweeps;
    int64_t warmup_sweeps;                   /* caller computes round(n_sweeps*warmup_ratio), lib.rs:219-220 */
    int32_t sweep_mode;                      /* PP_SWEEP_* */
    int64_t pt_interval;                     /* 0 = None */
    int32_t pt_schedule;                     /* PP_PT_* */
    /* 0 = None; > 0: Fortuin-Kasteleyn cluster update (clusters/fk.rs, action "update") of every system after each
     * cluster_update_interval-th sweep (simulation/mod.rs:434-470); int8 layouts with unit couplings, else PP_ERR_UNSUPPORTED */
    int64_t cluster_update_interval;
    /* 0 = None; > 0: Houdayer isoenergetic cluster move, group size 2 (clusters/overlap.rs:146-339, action "update"), for every
     * temperature after each overlap_cluster_update_interval-th sweep's measurements (simulation/mod.rs:596-746); int8 layout,
     * n_replicas >= 2, else an error before any state mutation */
    int64_t overlap_cluster_update_interval;
    /* 0 = None; > 0: integrated autocorrelation times of m^2 and q^2 over the recorded sweeps (statistics/autocorrelation.rs,
     * ring backend; the lag is clamped to [1, recorded sweeps / 4] as simulation/mod.rs:342-344) */
    int64_t autocorrelation_max_lag;
    int64_t snapshot_interval;                /* unsupported: must be 0 */
    /* 1: equilibration diagnostic (statistics/equilibration.rs): energies and link overlaps are measured after EVERY sweep and
     * their running averages reported at the checkpoints of pp_equil_checkpoints(n_sweeps) */
    int32_t equilibration_diagnostic;
    /* 1: fp32-coupling / Gibbs log thresholds are read from a host-libm logf table so that spin
     * trajectories are bit-identical to the CPU rule; 0: device logf (results agree to tolerance) */
    int32_t exact_log;
    int32_t cluster_mode;                    /* PP_CLUSTER_SW (0) or PP_CLUSTER_WOLFF (1); used when cluster_update_interval > 0 */
    int32_t overlap_cluster_mode;            /* PP_CLUSTER_*; used when overlap_cluster_update_interval > 0 (reference default: wolff) */
} pp_sample_cfg;

/* Result buffers; every pointer may be NULL (that output is skipped).
 * Shapes use T = n_temps, R = n_replicas, D = n_disorder of this handle, N = n_spins. */
typedef struct {
    double *mags, *mags2, *mags4, *energies, *energies2;                  /* [T]: mean over this handle's realizations */
    double *overlap, *overlap2, *overlap4;                                /* [T], R >= 2 */
    double *link_overlap, *link_overlap2, *link_overlap4;                 /* [T], R >= 2 */
    uint64_t *overlap_histogram;                                          /* [T][N+1] summed over realizations */
    double *ql_at_q_sum, *ql2_at_q_sum;                                   /* [T][N+1] summed over realizations */
    uint64_t *per_sample_overlap_histogram;                               /* [D][T][N+1] */
    double *per_sample_ql_at_q_sum, *per_sample_ql2_at_q_sum;             /* [D][T][N+1] */
    uint64_t *pt_edge_attempts, *pt_edge_acceptances;                     /* [D][T-1] */
    uint64_t *pt_round_trips;                                             /* [D][R][T] */
    double *per_sample_means;  /* [D][11][T]: per-realization averages in the order mags, mags2, mags4, energies,
                                  energies2, overlap, overlap2, overlap4, link_overlap, link_overlap2, link_overlap4;
                                  lets a multi-GPU caller do the reference's ordered sum over realizations */
    /* cfg.autocorrelation_max_lag > 0 (src/lib.rs:545-556): [T] means over this handle's realizations of the per-realization
     * Sokal-windowed taus (statistics/results.rs:217-231, 269-274); overlap2_tau needs R >= 2 */
    double *mags2_tau, *overlap2_tau;
    double *per_sample_taus;   /* [D][2][T]: per-realization taus (m^2 row, q^2 row), for the ordered multi-GPU merge */
    /* cfg.equilibration_diagnostic (src/lib.rs:559-574): [n_ckpt][T] running averages at the checkpoints, mean over this
     * handle's realizations (statistics/results.rs:231-247, 275-282) */
    double *equil_energy_avg, *equil_link_overlap_avg;
    double *per_sample_equil;  /* [D][n_ckpt][2][T] per-realization checkpoints (energy row, link-overlap row) */
} pp_results;

const char *pp_last_error(void);
int32_t pp_abi_version(void);
/* sizeof(pp_model_desc / pp_sample_cfg / pp_results) as this library was compiled (which = 0 / 1 / 2): lets a binding that declares
 * the structs by hand (ctypes, a Rust #[repr(C)] block) check its layout at load time */
int64_t pp_struct_size(int32_t which);

/* host-only helpers (no GPU needed) */
pp_status pp_colouring(int32_t n_dims, const int64_t *shape, int32_t n_offsets, const int64_t *offsets,
                       uint16_t *colour_out /* [n_spins] */, int32_t *n_colours_out);
pp_status pp_metropolis_lookup(const float *temperatures, int32_t n_temps, int32_t n_neighbors, int32_t sweep_mode,
                               uint32_t *table_out /* [n_temps][4*n_neighbors+1] */);
uint64_t pp_realization_seed(uint64_t root, uint64_t realization);
/* statistics/equilibration.rs:18-29: 128, 256, ... < n_sweeps, then n_sweeps; returns the count, fills out[0..count) if not NULL */
int32_t pp_equil_checkpoints(int64_t n_sweeps, int64_t *out);

/* state-owning engine */
pp_status pp_create(const pp_model_desc *desc, pp_sim **out);
void pp_destroy(pp_sim *sim);
pp_status pp_sample(pp_sim *sim, const pp_sample_cfg *cfg, pp_results *out, const volatile int32_t *interrupt,
                    void (*on_sweep)(void *user, uint64_t sweep_id), void *user);
pp_status pp_reset(pp_sim *sim, int32_t has_seed, uint64_t seed);
/* slab layout: [S][local planes][L1][L2] of this process's slab (the whole lattice when slab_rank = -1 or slab_ranks <= 1) */
pp_status pp_get_spins(pp_sim *sim, int64_t realization, int8_t *out /* [S*N], system-major */);
/* spins per system held by this handle (N, or N / slab_ranks for one slab of a decomposed lattice) */
int64_t pp_local_spin_count(const pp_sim *sim);
/* NCCL bootstrap for PP_LAYOUT_SLAB across processes: rank 0 calls this, the host broadcasts the bytes */
pp_status pp_nccl_unique_id(uint8_t *out /* [PP_NCCL_ID_BYTES] */);
/* 1 when this process already holds the communicator of (device, world size, rank): handles created from now on reuse it and
 * need no bootstrap token (every rank of the world must then pass nccl_unique_id = NULL, or every rank a fresh token) */
int32_t pp_nccl_comm_cached(int32_t device, int32_t ranks, int32_t rank);
pp_status pp_get_system_ids(pp_sim *sim, int64_t realization, int64_t *out /* [S] */);
pp_status pp_get_energies(pp_sim *sim, int64_t realization, float *out /* [S] by system */);
int32_t pp_get_layout(const pp_sim *sim);
/* Measurement hooks (bench.py, tools/): not part of the reference boundary, kept out of the production structs.
 * pp_debug_set_profile(sim, 1): the NEXT pp_sample call brackets every sweep-kernel launch with CUDA events on the launch stream
 * (one stream, one launch per sweep) — the roofline figure's kernel time.  pp_debug_last_timing: figures of the last pp_sample. */
typedef struct {
    double sweep_loop_ms;          /* device time of the sweep loop (CUDA events on the launch stream) */
    int64_t kernel_launches;       /* kernels (and collectives) launched by the call */
    double sweep_kernel_ms;        /* profile mode: summed device time of the sweep-kernel launches */
    int64_t sweep_kernel_launches; /* profile mode: how many launches that sum covers */
} pp_timing;
pp_status pp_debug_set_profile(pp_sim *sim, int32_t on);
pp_status pp_debug_last_timing(const pp_sim *sim, pp_timing *out);
/* 1 when sweeps run through the stride-based 3-D multispin kernel (pp_kernels_msc3d.cuh), else 0 */
int32_t pp_uses_msc3d(const pp_sim *sim);
/* 1 when a PP_LAYOUT_SLAB handle stores one bit per spin (shape[2] % 64 == 0: pp_kernels_slabp.cuh, packed draw mapping), else 0 */
int32_t pp_slab_packed(const pp_sim *sim);
/* 1 when an int8-layout handle keeps its ferromagnetic systems as one bit per spin, resident in shared memory per launch
 * (pp_kernels_prows.cuh: row-alternating colouring, last extent % 64 == 0; packed draw mapping), else 0 */
int32_t pp_rows_packed(const pp_sim *sim);
/* 1 when an int8-layout handle with fp32 couplings keeps the same site of 32 systems of a realization in one word, one bit per spin
 * (pp_kernels_swords.cuh: two-colour row-alternating lattice, last extent % 32 == 0 or 8 / 16 / 24 with N % 32 == 0, >= 16 systems; system-quad draw mapping), else 0 */
int32_t pp_sys_words(const pp_sim *sim);

/* operator-level entry points with the reference's slice semantics (unit-level parity tests):
 * H2D -> kernel -> D2H on the handle's state. */
pp_status pp_set_spins(pp_sim *sim, int64_t realization, const int8_t *spins /* [S*N] */);
pp_status pp_set_system_ids(pp_sim *sim, int64_t realization, const int64_t *ids /* [S] */);
pp_status pp_op_sweep(pp_sim *sim, int32_t sweep_mode, uint32_t sweep_index, int32_t exact_log);
pp_status pp_op_energies_mags(pp_sim *sim, float *energies /* [D][S] */, int64_t *mags /* [D][S] */);
pp_status pp_op_overlap(pp_sim *sim, int64_t *dot_spin /* [D][P][T] */, int64_t *dot_link /* [D][P][T] */);
pp_status pp_op_pt(pp_sim *sim, int32_t pt_schedule, uint32_t pt_event);

/* The same operators with the reference's HOST-SLICE signatures: one realization (model->n_disorder = 1), the
 * slices go to the device, one kernel runs, the result comes back; a temporary handle lives for the call.
 *   pp_slice_sweep          <- metropolis_sweep / gibbs_sweep(lattice, spins, couplings, temperatures, system_ids, rngs, ..)
 *                              spin-sim/src/mcmc/sweep.rs:220-229, 262-270 (draws: RNG-SPEC with model->seed and sweep_index)
 *   pp_slice_energies_mags  <- compute_energies_and_magnetizations_into(lattice, spins, couplings, energies, mags)
 *                              spin-sim/src/spins/energy.rs:59-65
 *   pp_slice_overlap        <- OverlapAccum::collect(lattice, spins, system_ids, ..)   spin-sim/src/statistics/overlap.rs:251
 *   pp_slice_pt             <- parallel_tempering(_full_ladder)(energies, temperatures, system_ids, n_spins, rng, ..)
 *                              spin-sim/src/mcmc/tempering.rs:20-27, 45-53 */
pp_status pp_slice_sweep(const pp_model_desc *model, int32_t sweep_mode, uint32_t sweep_index, int32_t exact_log,
                         int8_t *spins /* [S*N] system-major, in/out */, const int64_t *system_ids /* [S] slot -> system, or NULL */);
pp_status pp_slice_energies_mags(const pp_model_desc *model, const int8_t *spins /* [S*N] */, float *energies /* [S] */,
                                 int64_t *mags /* [S] or NULL */);
pp_status pp_slice_overlap(const pp_model_desc *model, const int8_t *spins /* [S*N] */, const int64_t *system_ids /* [S] or NULL */,
                           int64_t *dot_spin /* [P][T] */, int64_t *dot_link /* [P][T] */);
pp_status pp_slice_pt(const pp_model_desc *model, int32_t pt_schedule, uint32_t pt_event, int32_t first_parity,
                      const float *energies /* [S] by system */, int64_t *system_ids /* [S] in/out */);

#ifdef __cplusplus
}
#endif
#endif
