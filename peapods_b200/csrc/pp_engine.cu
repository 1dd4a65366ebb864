// pp_engine.cu — C ABI + host driver of the B200 sweep engine (see include/peapods_b200.h).
//
// The host driver restates the sequencing of run_sweep_loop_impl
// (spin-sim/src/simulation/mod.rs:405-432, 486-509, 527-529, 543-578, 748-796): per sweep
//   sweep -> energies(+mags) if record||pt -> overlap (pre-swap system_ids) if record
//         -> fold if record -> parallel tempering if pt_this_sweep
// and enqueues one kernel per step on a single stream; no spin data crosses PCIe during sample().
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/peapods_b200.h"
#include "pp_device.cuh"
#include "pp_kernels_fk.cuh"
#include "pp_kernels_int8.cuh"
#include "pp_kernels_msc.cuh"
#include "pp_kernels_msc3d.cuh"
#include "pp_kernels_rows.cuh"
#include "pp_kernels_prows.cuh"
#include "pp_kernels_swords.cuh"
#include "pp_kernels_stats.cuh"
#include "pp_plan.h"
#include "pp_slab.cuh"

using namespace pp;

// ------------------------------------------------------------------------------------------
// errors
static thread_local std::string g_last_error;
static pp_status fail(pp_status st, const std::string &msg) {
    g_last_error = msg;
    return st;
}
extern "C" const char *pp_last_error(void) { return g_last_error.c_str(); }
extern "C" int32_t pp_abi_version(void) { return PP_ABI_VERSION; }
extern "C" int64_t pp_struct_size(int32_t which) {
    switch (which) {
        case 0: return (int64_t)sizeof(pp_model_desc);
        case 1: return (int64_t)sizeof(pp_sample_cfg);
        case 2: return (int64_t)sizeof(pp_results);
    }
    return -1;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(_e == cudaErrorMemoryAllocation ? PP_ERR_OOM : PP_ERR_CUDA,                  \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
    } while (0)

// ------------------------------------------------------------------------------------------
// host-side acceptance tables (mcmc/sweep.rs:141-165 restated with the host libm logf)
static const uint32_t F24 = 1u << 24;

static bool metropolis_accepts(float temperature, int32_t ec, uint32_t draw) {  // sweep.rs:161-165
    volatile float uniform = (float)draw / (float)F24;
    volatile float half_t = temperature / 2.0f;
    volatile float rhs = half_t * logf(uniform);
    return (float)ec >= rhs;
}
static bool gibbs_accepts(float temperature, int32_t ec, uint32_t draw) {  // sweep.rs:279-282 on the 24-bit grid
    volatile float u = (float)draw / (float)F24;
    volatile float one_minus = 1.0f - u;
    volatile float ratio = u / one_minus;
    volatile float half_t = temperature / 2.0f;
    volatile float rhs = half_t * logf(ratio);
    return (float)ec >= rhs;
}
static uint32_t accepted_count(float temperature, int32_t ec, bool gibbs) {  // sweep.rs:147-159
    uint32_t low = 0, high = F24;
    while (low < high) {
        uint32_t mid = low + (high - low) / 2;
        bool acc = gibbs ? gibbs_accepts(temperature, ec, mid) : metropolis_accepts(temperature, ec, mid);
        if (acc) low = mid + 1;
        else high = mid;
    }
    return low;
}
static bool temps_eligible(const float *t, int n) {  // sweep.rs:114-117
    for (int i = 0; i < n; i++)
        if (!(std::isfinite(t[i]) && t[i] / 2.0f > 0.0f)) return false;
    return true;
}

extern "C" pp_status pp_metropolis_lookup(const float *temperatures, int32_t n_temps, int32_t n_neighbors,
                                          int32_t sweep_mode, uint32_t *table_out) {
    if (!temperatures || !table_out || n_temps < 0 || n_neighbors < 1) return fail(PP_ERR_INVALID, "bad arguments");
    if (!temps_eligible(temperatures, n_temps))
        return fail(PP_ERR_INVALID, "temperatures must be finite with T/2 > 0 for the integer lookup");
    const int off = 2 * n_neighbors, width = 4 * n_neighbors + 1;
    for (int t = 0; t < n_temps; t++)
        for (int ec = -off; ec <= off; ec++)
            table_out[t * width + ec + off] = accepted_count(temperatures[t], ec, sweep_mode == PP_SWEEP_GIBBS);
    return PP_OK;
}

extern "C" uint64_t pp_realization_seed(uint64_t root, uint64_t r) { return realization_seed(root, r); }

extern "C" int32_t pp_equil_checkpoints(int64_t n_sweeps, int64_t *out) {  // equilibration.rs:18-29
    int32_t n = 0;
    int64_t last = -1;
    for (int64_t p = 128; p < n_sweeps; p *= 2) {
        if (out) out[n] = p;
        last = p;
        n++;
    }
    if (last != n_sweeps) {
        if (out) out[n] = n_sweeps;
        n++;
    }
    return n;
}

extern "C" pp_status pp_colouring(int32_t n_dims, const int64_t *shape, int32_t n_offsets, const int64_t *offsets,
                                  uint16_t *colour_out, int32_t *n_colours_out) {
    if (!shape) return fail(PP_ERR_INVALID, "shape is NULL");
    LatticePlan plan;
    std::string err = build_plan(n_dims, shape, n_offsets, offsets, plan);
    if (!err.empty()) return fail(PP_ERR_UNSUPPORTED, err);
    if (colour_out) memcpy(colour_out, plan.colour.data(), sizeof(uint16_t) * (size_t)plan.n_spins);
    if (n_colours_out) *n_colours_out = plan.n_colours;
    return PP_OK;
}

// ------------------------------------------------------------------------------------------
// host-libm log tables (bit-exact thresholds): logtab[d] = logf(d/2^24), glogtab[d] = logf(u/(1-u))
struct LogTables {
    float *logtab = nullptr, *glogtab = nullptr;
};
static std::mutex g_tab_mutex;
static std::map<int, LogTables> g_tabs;

static pp_status get_log_table(int device, bool gibbs, const float **out) {
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    LogTables &lt = g_tabs[device];
    float *&slot = gibbs ? lt.glogtab : lt.logtab;
    if (!slot) {
        std::vector<float> host(F24);
        unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (unsigned w = 0; w < nthreads; w++)
            pool.emplace_back([&, w]() {
                for (uint32_t d = w; d < F24; d += nthreads) {
                    volatile float u = (float)d / (float)F24;
                    if (gibbs) {
                        volatile float om = 1.0f - u;
                        volatile float ratio = u / om;
                        host[d] = logf(ratio);
                    } else {
                        host[d] = logf(u);
                    }
                }
            });
        for (auto &th : pool) th.join();
        CUDA_TRY(cudaMalloc((void **)&slot, sizeof(float) * (size_t)F24)); /*static*/
        CUDA_TRY(cudaMemcpy(slot, host.data(), sizeof(float) * (size_t)F24, cudaMemcpyHostToDevice));
    }
    *out = slot;
    return PP_OK;
}

// ------------------------------------------------------------------------------------------
struct pp_sim {
    LatticePlan plan;
    ModelView mv{};
    int layout = PP_LAYOUT_INT8;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<float> temps;
    uint64_t ctor_seed = 0;
    uint32_t sweep_counter = 0, pt_event_counter = 0;  // RNG-SPEC counters; persist across sample()
    int next_parity = 0;                               // realization.rs:84-90
    int64_t G = 0;                                     // word groups (MSC)
    bool msc3d = false;                                // specialised 3-D hypercubic MSC kernel usable
    bool msc3d_metro = false;                          // Metropolis counts for unsat >= 3 are all 2^24
    int msc3d_nh = 2;                                  // temperature slots (256-thread halves) per CTA of the msc3d kernel
    size_t msc3d_smem = 0;
    bool msc3d_esw = false;                            // in-sweep energy counters + pair-only epilogue (see msc3d_kernel)
    Msc3dPlan m3;
    Msc3dView gv{};
    uint32_t *d_perm = nullptr;
    uint4 *d_items = nullptr;
    // owned device buffers
    uint32_t *d_nbr = nullptr, *d_order = nullptr, *d_colour_start = nullptr;
    int8_t *d_J8 = nullptr;
    float *d_Jf = nullptr;
    uint32_t *d_Jw = nullptr;
    int8_t *d_spins = nullptr;
    uint32_t *d_words = nullptr, *d_words_alt = nullptr;
    int32_t *d_sid = nullptr;
    float *d_energies = nullptr, *d_temps = nullptr;
    long long *d_mags = nullptr;
    uint32_t *d_lut_metro = nullptr, *d_lut_gibbs = nullptr;
    // PT
    PtView pt{};
    // stats
    StatsView st{};
    long long *d_dot_spin = nullptr, *d_dot_link = nullptr;
    bool hist_allocated = false;
    int64_t launches = 0;
    // chunked multi-stream execution of the msc3d path (see pp_sample)
    int n_streams = 8;
    int64_t chunk_bytes = 0;                           // 0: one chunk per stream, between 8 and 32 MiB of spin words (measured best)
    int64_t macro_batch = 16;
    int64_t chunk_groups = 0;                          // > 0: word groups per chunk (overrides chunk_bytes; tests)
    int64_t max_batch = 64;                            // sweeps without reduction / PT fused into one launch (multispin)
    bool defer_swaps = true;                           // gather PT lane swaps in the next sweep's stage-in
    std::vector<cudaStream_t> xstreams;
    std::vector<cudaEvent_t> xevents;
    SlabState *slab = nullptr;                         // PP_LAYOUT_SLAB (pp_slab.cuh)
    // system-split handle (pp_model_desc.system_ranks > 1): this process sweeps the systems [mv.sys_lo, mv.sys_hi) of the one
    // realization; energies / magnetisations / configurations are all-gathered over sys_comm (cached, never destroyed here)
    int sys_ranks = 1, sys_rank = 0;
    ncclComm_t sys_comm = nullptr;
    bool rows = false;                                 // int8 layout through the per-row stride tables (pp_kernels_rows.cuh)
    bool resident = false;                             // small realizations: one CTA per realization, many sweeps per launch
    uint16_t *d_site16 = nullptr;                      // storage index -> logical site (same use)
    uint16_t *d_nbr16 = nullptr;                       // storage-space neighbour table as u16 (multispin Houdayer move), built on first use
    bool rows_esw = false;                             // two-colour lattice: the last colour pass also delivers the energies
    float rows_escale = 1.0f;                          // fp32 couplings: fixed-point unit of the in-sweep bond sums (power of two)
    size_t resident_smem = 0;
    RowsView rv{};
    std::vector<void *> rows_bufs;
    std::vector<uint32_t> rows_class_start;
    uint64_t *d_keys = nullptr;
    // ferromagnets with one bit per spin (pp_kernels_prows.cuh): the packed words are the state, `d_spins` is an int8 scratch view
    // that the API and the int8-only kernels (cluster moves) see through prows_sync(): unpack before, pack after
    bool prows = false;
    bool resident_packed = false;                      // small ferromagnetic realizations: prows_resident_kernel instead of rows_resident_kernel
    int resident_cluster = 0;                          // > 0: prows_cluster_resident_kernel, the systems of a realization over this many CTAs
    int resident_cluster_threads = 0;
    size_t resident_packed_smem = 0;
    PRowsView pv{};
    int prows_nm[2] = {0, 0};                          // thresholds compared per site: [metropolis, gibbs]
    bool prows_cluster = false;                        // recorded sweeps folded inside the sweep kernel by clusters of R CTAs (launch_prows)
    bool prows_bs[2] = {false, false};                 // z' = 3, all seven counts below 2^24 and growing with unsat: binary-search form
    // fp32 couplings with the same site of 32 systems in one word (pp_kernels_swords.cuh); `d_spins` is a scratch view as above
    bool swords = false;
    bool sw_tbits_valid = false;                       // the transposed view matches the words
    SWordsView swv{};
    long long *d_rows_acc = nullptr;                   // [2 * max(D*S, D*P*T)] split-reduction scratch (kept zero between launches)
    unsigned int *d_rows_arrive = nullptr;
    int rows_nb = 1;                                   // blocks per system / pair of the split reductions
    // measurement hook: event pairs around sweep-kernel launches
    bool profile = false, profile_next = false;        // pp_debug_set_profile arms the next pp_sample
    pp_timing last_timing{};
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
};

static void prof_mark(pp_sim *s, cudaStream_t stream) {
    if (!s->profile) return;
    if (s->prof_used == s->prof_events.size()) {
        if (s->prof_events.size() >= 8192) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        s->prof_events.push_back(e);
    }
    cudaEventRecord(s->prof_events[s->prof_used++], stream);
}

// Device buffers come from the device's default stream-ordered pool (unlimited release threshold) through an exact-size
// block cache: the reference API builds one IsingSimulation per model, and a destroyed handle's buffers are handed to
// the next handle of the same shape without touching the driver (the pool alone re-maps physical memory whenever its
// best-fit carving of the freed 4 GiB histogram blocks leaves no hole large enough: 60-500 ms per construction).
// A block enters the cache only after the work that used it has completed (callers synchronise first).
struct BlockCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> free_blocks;  // (device, bytes) -> block
    std::map<void *, std::pair<int, size_t>> live;              // block -> (device, bytes)
    size_t cached_bytes = 0, cap_bytes = size_t(64) << 30;
    BlockCache() {
        if (const char *e = getenv("PP_CACHE_GIB")) cap_bytes = (size_t)std::max(0, atoi(e)) << 30;
    }
};
static BlockCache &block_cache() {
    static BlockCache *c = new BlockCache();  // never destroyed: handles may outlive static destructors
    return *c;
}

static cudaError_t pool_alloc(pp_sim *s, void **p, size_t bytes) {
    static std::mutex mu;
    static std::map<int, bool> configured;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[s->device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, s->device) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            configured[s->device] = true;
        }
    }
    if (bytes == 0) bytes = 1;
    BlockCache &c = block_cache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.free_blocks.find({s->device, bytes});
        if (it != c.free_blocks.end()) {
            *p = it->second;
            c.free_blocks.erase(it);
            c.cached_bytes -= bytes;
            c.live[*p] = {s->device, bytes};
            return cudaSuccess;
        }
    }
    cudaError_t err = cudaMallocAsync(p, bytes, s->stream);
    if (err != cudaSuccess) {  // out of memory with blocks parked in the cache: release them and retry once
        cudaGetLastError();
        std::vector<void *> drop;
        {
            std::lock_guard<std::mutex> lock(c.mu);
            for (auto it = c.free_blocks.begin(); it != c.free_blocks.end();) {
                if (it->first.first == s->device) {
                    drop.push_back(it->second);
                    c.cached_bytes -= it->first.second;
                    it = c.free_blocks.erase(it);
                } else ++it;
            }
        }
        for (void *b : drop) cudaFreeAsync(b, s->stream);
        cudaStreamSynchronize(s->stream);
        err = cudaMallocAsync(p, bytes, s->stream);
    }
    if (err == cudaSuccess) {
        std::lock_guard<std::mutex> lock(c.mu);
        c.live[*p] = {s->device, bytes};
    }
    return err;
}
// the caller guarantees that no enqueued work still uses p
static void pool_free(pp_sim *s, void *p) {
    if (!p) return;
    BlockCache &c = block_cache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) {
            const std::pair<int, size_t> key = it->second;
            c.live.erase(it);
            if (c.cached_bytes + key.second <= c.cap_bytes) {
                c.free_blocks.insert({key, p});
                c.cached_bytes += key.second;
                return;
            }
        }
    }
    cudaFreeAsync(p, s->stream);
}

static void free_sim(pp_sim *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    void *ptrs[] = {s->d_perm, s->d_items, s->d_nbr, s->d_order, s->d_colour_start, s->d_J8, s->d_Jf, s->d_Jw, s->d_spins, s->d_words,
                    s->d_words_alt, s->d_sid, s->d_energies, s->d_temps, s->d_mags, s->d_lut_metro, s->d_lut_gibbs,
                    s->pt.edge_attempts, s->pt.edge_acceptances, s->pt.round_trips, s->pt.trip_state, s->pt.swap_mask,
                    s->st.sums, s->st.hist, s->st.ql_at_q, s->st.ql2_at_q, s->d_dot_spin, s->d_dot_link};
    if (s->stream) cudaStreamSynchronize(s->stream);  // pp_sample joins its side streams into this one before it returns
    for (cudaStream_t x : s->xstreams) cudaStreamSynchronize(x);
    for (void *p : ptrs)
        if (p) pool_free(s, p);
    for (void *b : s->rows_bufs) pool_free(s, b);
    if (s->d_keys) pool_free(s, s->d_keys);
    if (s->d_nbr16) pool_free(s, s->d_nbr16);
    if (s->d_site16) pool_free(s, s->d_site16);
    if (s->pv.words) pool_free(s, s->pv.words);
    if (s->d_rows_acc) pool_free(s, s->d_rows_acc);
    if (s->d_rows_arrive) pool_free(s, s->d_rows_arrive);
    if (s->slab) {
        SlabState *sl = s->slab;
        if (sl->comm_stream) cudaStreamSynchronize(sl->comm_stream);
        if (sl->comm && !sl->comm_cached) nccl_api().CommDestroy(sl->comm);
        for (uint8_t *b : sl->buffers) cudaFree(b);
        for (SlabPView &v : sl->pparts) cudaFree(v.words);
        if (sl->ev_main) cudaEventDestroy(sl->ev_main);
        if (sl->d_partial) cudaFree(sl->d_partial);
        if (sl->ev_boundary) cudaEventDestroy(sl->ev_boundary);
        if (sl->ev_halo) cudaEventDestroy(sl->ev_halo);
        if (sl->comm_stream) cudaStreamDestroy(sl->comm_stream);
        delete sl;
    }
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : s->xevents) cudaEventDestroy(e);
    for (cudaStream_t x : s->xstreams) cudaStreamDestroy(x);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" void pp_destroy(pp_sim *sim) { free_sim(sim); }
extern "C" int32_t pp_get_layout(const pp_sim *sim) { return sim ? sim->layout : 0; }
extern "C" pp_status pp_debug_set_profile(pp_sim *sim, int32_t on) {
    if (!sim) return fail(PP_ERR_INVALID, "sim is NULL");
    sim->profile_next = on != 0;
    return PP_OK;
}
extern "C" pp_status pp_debug_last_timing(const pp_sim *sim, pp_timing *out) {
    if (!sim || !out) return fail(PP_ERR_INVALID, "sim/out is NULL");
    *out = sim->last_timing;
    return PP_OK;
}
#ifdef PP_M3_TIMING
extern "C" int32_t pp_debug_m3_clocks(unsigned long long *out, int64_t n) {  // variant builds only (tools/m3_phases.py)
    cudaDeviceSynchronize();
    return (int32_t)cudaMemcpyFromSymbol(out, pp::pp_m3_clk, (size_t)n * sizeof(unsigned long long));
}
#endif
extern "C" int32_t pp_debug_prows_clocks(unsigned long long *out, int32_t reset) {  // variant builds only (tools/c3_phases.py)
#ifdef PP_PROWS_TIMING
    if (reset) {
        unsigned long long z[8] = {};
        return cudaMemcpyToSymbol(pp::pp_prows_clk, z, sizeof(z)) == cudaSuccess ? 1 : 0;
    }
    return cudaMemcpyFromSymbol(out, pp::pp_prows_clk, sizeof(unsigned long long) * 8) == cudaSuccess ? 1 : 0;
#else
    (void)out; (void)reset;
    return 0;
#endif
}
extern "C" int32_t pp_uses_msc3d(const pp_sim *sim) { return sim && sim->msc3d ? 1 : 0; }
extern "C" int32_t pp_slab_packed(const pp_sim *sim) { return sim && sim->slab && sim->slab->packed ? 1 : 0; }
extern "C" int32_t pp_sys_words(const pp_sim *sim) { return sim && sim->swords ? 1 : 0; }
extern "C" int32_t pp_rows_packed(const pp_sim *sim) { return sim && (sim->prows || sim->rv.packed_draws) ? 1 : 0; }
extern "C" int64_t pp_local_spin_count(const pp_sim *sim) {
    if (!sim) return 0;
    return sim->slab ? sim->slab->local_planes() * sim->slab->plane : sim->mv.N;
}
extern "C" pp_status pp_nccl_unique_id(uint8_t *out) {
    if (!out) return fail(PP_ERR_INVALID, "out is NULL");
    NcclApi &nc = nccl_api();
    if (!nc.error.empty()) return fail(PP_ERR_NCCL, nc.error);
    ncclUniqueId id;
    ncclResult_t r = nc.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(PP_ERR_NCCL, std::string("ncclGetUniqueId: ") + nc.GetErrorString(r));
    memcpy(out, &id, PP_NCCL_ID_BYTES);
    return PP_OK;
}

// coupling classification: flags[0] non-unit value, flags[1] zero, flags[2] negative
__global__ void classify_couplings_kernel(const float *J, int64_t n, int *flags) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f0 = 0, f1 = 0, f2 = 0;
    for (int64_t i = gid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float c = J[i];
        if (!(c == -1.0f || c == 0.0f || c == 1.0f)) f0 = 1;  // sweep.rs:110-112 (NaN fails closed)
        if (c == 0.0f) f1 = 1;
        if (c < 0.0f) f2 = 1;
    }
    if (f0) atomicOr(&flags[0], 1);
    if (f1) atomicOr(&flags[1], 1);
    if (f2) atomicOr(&flags[2], 1);
}
__global__ void couplings_to_int8_kernel(const float *J, int8_t *out, int64_t n) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) out[gid] = (int8_t)J[gid];
}
__global__ void iota_sid_kernel(int32_t *sid, int64_t n, int S) {
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) sid[gid] = (int32_t)(gid % S);
}
__global__ void pt_mark_hot_kernel(ModelView m, PtView pt) {  // realization.rs:64-66
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= m.D * m.R) return;
    int64_t d = gid / m.R;
    int r = (int)(gid % m.R);
    pt.trip_state[d * m.S + m.system_ids[d * m.S + r * m.T + pt.hot_slot]] = 1;
}

static inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- kernel launch helpers ------------------------------------------------------------------
// A Ctx is a view of a contiguous block of realizations (a "chunk": whole word groups for the multispin layout) plus
// the stream its kernels go to.  All device arrays are indexed by realization / word group, so a chunk is the same
// structs with advanced base pointers; seeds use sample_offset + d, so results do not depend on the chunking.
struct Ctx {
    ModelView m;
    StatsView st;
    PtView pt;
    long long *dot_spin, *dot_link;
    int64_t G;
    cudaStream_t stream;
    // msc3d ping-pong / deferred parallel-tempering swaps (see msc3d_kernel): m.words is the current buffer
    uint32_t *words_alt = nullptr;
    int flips = 0;
    bool swap_pending = false;
    int pend_schedule = 0, pend_parity = 0;
};

static Ctx whole_ctx(pp_sim *s) {
    Ctx c;
    c.words_alt = s->d_words_alt;
    c.m = s->mv;
    c.st = s->st;
    c.pt = s->pt;
    c.dot_spin = s->d_dot_spin;
    c.dot_link = s->d_dot_link;
    c.G = s->G;
    c.stream = s->stream;
    return c;
}

// realizations [d0, d0 + D) of the handle; for the multispin layout d0 is a multiple of 32
static Ctx chunk_ctx(pp_sim *s, int64_t d0, int64_t D, cudaStream_t stream) {
    Ctx c = whole_ctx(s);
    ModelView &m = c.m;
    const int64_t N = m.N, S = m.S, T = m.T, g0 = d0 / 32;
    m.D = D;
    m.sample_offset = s->mv.sample_offset + d0;
    if (m.J8) m.J8 += d0 * N * m.z;
    if (m.Jf) m.Jf += d0 * N * m.z;
    if (m.Jw) m.Jw += g0 * m.z * N;
    if (m.spins) m.spins += d0 * S * N;
    if (m.words) m.words += g0 * S * N;
    if (c.words_alt) c.words_alt += g0 * S * N;
    m.system_ids += d0 * S;
    m.energies += d0 * S;
    m.mags += d0 * S;
    c.st.sums += d0 * 11 * T;
    if (c.st.hist) {
        c.st.hist += d0 * T * (N + 1);
        c.st.ql_at_q += d0 * T * (N + 1);
        c.st.ql2_at_q += d0 * T * (N + 1);
    }
    if (c.dot_spin) {
        c.dot_spin += d0 * m.P * T;
        c.dot_link += d0 * m.P * T;
        c.st.dot_spin = c.dot_spin;
        c.st.dot_link = c.dot_link;
    }
    c.pt.edge_attempts += d0 * (T > 1 ? T - 1 : 1);
    c.pt.edge_acceptances += d0 * (T > 1 ? T - 1 : 1);
    c.pt.round_trips += d0 * S;
    c.pt.trip_state += d0 * S;
    if (c.pt.swap_mask) c.pt.swap_mask += g0 * m.R * (T > 1 ? T - 1 : 1);
    c.G = (D + 31) / 32;
    c.stream = stream;
    return c;
}

static pp_status launch_energy(pp_sim *s, Ctx &c, bool want_mags);

template <int RPC, bool METRO, int NH, int NFIX = 0>
static pp_status launch_msc3d_t(pp_sim *s, Ctx &c, const ModelView &m, uint32_t sweep_index, int n_sweeps,
                                bool want_energy, bool want_mags, bool want_overlap, bool want_fold) {
    static bool configured[64] = {};  // per instantiation and device (the attribute is a per-device property of the function)
    if (s->device < 0 || s->device >= 64 || !configured[s->device]) {
        CUDA_TRY(cudaFuncSetAttribute(msc3d_kernel<RPC, METRO, NH, NFIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (s->device >= 0 && s->device < 64) configured[s->device] = true;
    }
    const unsigned grid = (unsigned)(c.G * ((m.T + NH - 1) / NH));
    msc3d_kernel<RPC, METRO, NH, NFIX><<<grid, MSC3D_NTH * NH, s->msc3d_smem, c.stream>>>(
        m, s->gv, c.st, sweep_index, n_sweeps, want_energy, want_mags, want_overlap, want_fold, m.sample_offset / 32, c.dot_spin,
        c.dot_link, c.words_alt, c.swap_pending ? c.pt.swap_mask : nullptr, c.pend_schedule, c.pend_parity, s->msc3d_esw ? 1 : 0);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    if (n_sweeps > 0) {  // the sweep consumed the pending exchange and wrote the other buffer
        std::swap(c.m.words, c.words_alt);
        c.flips++;
        c.swap_pending = false;
    }
    return PP_OK;
}

// make the handle's current-buffer pointer follow the launches issued through a context
static void commit_ctx(pp_sim *s, const Ctx &c) {
    if (c.flips & 1) {
        std::swap(s->d_words, s->d_words_alt);
        s->mv.words = s->d_words;
    }
}

static pp_status launch_msc3d(pp_sim *s, Ctx &c, const ModelView &m, int sweep_mode, uint32_t sweep_index, int n_sweeps,
                              bool want_energy, bool want_mags, bool want_overlap, bool want_fold) {
    const bool metro = sweep_mode == PP_SWEEP_METROPOLIS && s->msc3d_metro;
    const bool two = s->msc3d_nh == 2;
#define PP_M3A(R_, M_, H_) launch_msc3d_t<R_, M_, H_>(s, c, m, sweep_index, n_sweeps, want_energy, want_mags, want_overlap, want_fold)
#define PP_M3(R_) return metro ? (two ? PP_M3A(R_, true, 2) : PP_M3A(R_, true, 1)) : (two ? PP_M3A(R_, false, 2) : PP_M3A(R_, false, 1))
    // the 16^3 lattice of the headline configuration with one slot per CTA: site count as a compile-time constant
    if (m.N == 4096 && !two && metro && (m.R == 2 || m.R == 4)) {
        if (m.R == 4) return launch_msc3d_t<4, true, 1, 4096>(s, c, m, sweep_index, n_sweeps, want_energy, want_mags, want_overlap, want_fold);
        return launch_msc3d_t<2, true, 1, 4096>(s, c, m, sweep_index, n_sweeps, want_energy, want_mags, want_overlap, want_fold);
    }
    switch (m.R) {
        case 1: PP_M3(1);
        case 2: PP_M3(2);
        case 4: PP_M3(4);
    }
#undef PP_M3
#undef PP_M3A
    return fail(PP_ERR_UNSUPPORTED, "msc3d: unsupported replica count");
}

// One communicator per (process, device, world, rank), kept for the life of the process: ncclCommInitRank costs seconds and a handle
// is built per model.  Every rank takes the same branch, so the bootstrap token of a later handle is simply not used.
static std::mutex comm_mu;
static std::map<std::tuple<int, int, int>, ncclComm_t> comm_cache;
extern "C" int32_t pp_nccl_comm_cached(int32_t device, int32_t ranks, int32_t rank) {
    std::lock_guard<std::mutex> lock(comm_mu);
    return comm_cache.count(std::tuple<int, int, int>{device, ranks, rank}) ? 1 : 0;
}
static pp_status cached_comm(int device, int ranks, int rank, const uint8_t *token, ncclComm_t *out) {
    NcclApi &nc = nccl_api();
    if (!nc.error.empty()) return fail(PP_ERR_NCCL, nc.error);
    std::lock_guard<std::mutex> lock(comm_mu);
    const std::tuple<int, int, int> ck{device, ranks, rank};
    auto it = comm_cache.find(ck);
    if (it != comm_cache.end()) {
        *out = it->second;
        return PP_OK;
    }
    if (!token) return fail(PP_ERR_INVALID, "nccl_unique_id is NULL");
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == PP_NCCL_ID_BYTES, "ncclUniqueId size");
    memcpy(&id, token, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = nc.CommInitRank(&comm, ranks, id, rank);
    if (r != ncclSuccess) return fail(PP_ERR_NCCL, std::string("ncclCommInitRank: ") + nc.GetErrorString(r));
    comm_cache[ck] = comm;
    *out = comm;
    return PP_OK;
}

// ---- slab layout (pp_slab.cuh) -----------------------------------------------------------------
#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) return fail(PP_ERR_NCCL, std::string(#expr) + ": " + nccl_api().GetErrorString(_r)); \
    } while (0)

// boundary planes of every local slab -> the halo planes of their neighbours, enqueued on `stream`.
// Bit-packed storage: colour = 0 / 1 sends only that colour's half of the planes (what a colour half-step changed), -1 both.
static pp_status slab_exchange(pp_sim *s, cudaStream_t stream, int colour = -1) {
    SlabState *sl = s->slab;
    const int S = s->mv.S;
    if (sl->packed) {
        const SlabPView &v0 = sl->pparts[0];
        const int64_t first = colour == 1 ? v0.half : 0, count = colour < 0 ? v0.plane : v0.half;  // words inside a plane
        const size_t bytes = sizeof(uint32_t) * (size_t)count;
        if (sl->rank < 0 || sl->ranks == 1) {
            const int n = (int)sl->pparts.size();
            for (int r = 0; r < n; r++) {
                const SlabPView &me = sl->pparts[(size_t)r], &lo = sl->pparts[(size_t)((r + n - 1) % n)], &up = sl->pparts[(size_t)((r + 1) % n)];
                for (int sys = 0; sys < S; sys++) {
                    const int64_t o = (int64_t)sys * me.sys_stride + first;
                    CUDA_TRY(cudaMemcpyAsync(lo.words + o + (int64_t)(sl->P + 1) * me.plane, me.words + o + me.plane, bytes,
                                             cudaMemcpyDeviceToDevice, stream));
                    CUDA_TRY(cudaMemcpyAsync(up.words + o, me.words + o + (int64_t)sl->P * me.plane, bytes, cudaMemcpyDeviceToDevice, stream));
                }
            }
            return PP_OK;
        }
        NcclApi &nc = nccl_api();
        const int lo = (sl->rank + sl->ranks - 1) % sl->ranks, up = (sl->rank + 1) % sl->ranks;
        NCCL_TRY(nc.GroupStart());
        for (int sys = 0; sys < S; sys++) {
            uint32_t *b = v0.words + (int64_t)sys * v0.sys_stride + first;
            NCCL_TRY(nc.Send(b + v0.plane, bytes, ncclUint8, lo, sl->comm, stream));
            NCCL_TRY(nc.Send(b + (int64_t)sl->P * v0.plane, bytes, ncclUint8, up, sl->comm, stream));
            NCCL_TRY(nc.Recv(b + (int64_t)(sl->P + 1) * v0.plane, bytes, ncclUint8, up, sl->comm, stream));
            NCCL_TRY(nc.Recv(b, bytes, ncclUint8, lo, sl->comm, stream));
        }
        NCCL_TRY(nc.GroupEnd());
        return PP_OK;
    }
    const size_t bytes = (size_t)sl->plane;
    if (sl->rank < 0 || sl->ranks == 1) {  // all slabs local: device copies
        const int n = (int)sl->parts.size();
        for (int r = 0; r < n; r++) {
            const SlabView &me = sl->parts[(size_t)r], &lo = sl->parts[(size_t)((r + n - 1) % n)], &up = sl->parts[(size_t)((r + 1) % n)];
            for (int sys = 0; sys < S; sys++) {
                const int64_t o = (int64_t)sys * me.sys_stride;
                CUDA_TRY(cudaMemcpyAsync(lo.spins + o + (int64_t)(sl->P + 1) * sl->plane, me.spins + o + sl->plane, bytes,
                                         cudaMemcpyDeviceToDevice, stream));
                CUDA_TRY(cudaMemcpyAsync(up.spins + o, me.spins + o + (int64_t)sl->P * sl->plane, bytes, cudaMemcpyDeviceToDevice, stream));
            }
        }
        return PP_OK;
    }
    NcclApi &nc = nccl_api();
    const SlabView &me = sl->parts[0];
    const int lo = (sl->rank + sl->ranks - 1) % sl->ranks, up = (sl->rank + 1) % sl->ranks;
    NCCL_TRY(nc.GroupStart());
    for (int sys = 0; sys < S; sys++) {
        uint8_t *b = me.spins + (int64_t)sys * me.sys_stride;
        NCCL_TRY(nc.Send(b + sl->plane, bytes, ncclUint8, lo, sl->comm, stream));                             // my first plane
        NCCL_TRY(nc.Send(b + (int64_t)sl->P * sl->plane, bytes, ncclUint8, up, sl->comm, stream));            // my last plane
        NCCL_TRY(nc.Recv(b + (int64_t)(sl->P + 1) * sl->plane, bytes, ncclUint8, up, sl->comm, stream));      // upper halo
        NCCL_TRY(nc.Recv(b, bytes, ncclUint8, lo, sl->comm, stream));                                         // lower halo
    }
    NCCL_TRY(nc.GroupEnd());
    return PP_OK;
}

// totals of the per-slab partial sums -> energies / magnetisations of every system, identical on every rank
static pp_status slab_finish_energy(pp_sim *s, const ModelView &m, cudaStream_t stream, bool want_mags) {
    SlabState *sl = s->slab;
    if (sl->comm)
        NCCL_TRY(nccl_api().AllReduce(sl->d_partial, sl->d_partial, 2 * (size_t)m.S, ncclUint64, ncclSum, sl->comm, stream));
    slab_finish_energy_kernel<<<blocks_for(m.S, 128), 128, 0, stream>>>(m, sl->d_partial, want_mags ? 1 : 0);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// one launch of the bit-packed kernel over planes [pa, pa + np) (+ plane pb when pb > 0)
static pp_status slabp_launch(pp_sim *s, const ModelView &m, const SlabPView &v, cudaStream_t stream, int colour, uint32_t sweep,
                              int pa, int np, int pb, int nm, bool update, bool acc) {
    const unsigned rows = (unsigned)(np + (pb > 0 ? 1 : 0));
    if (rows == 0) return PP_OK;
    const dim3 grid(blocks_for(v.half, SLABP_THREADS), rows, (unsigned)m.S);
    unsigned long long *partial = s->slab->d_partial;
#define PP_SLABP(NM_, U_, A_) slabp_sweep_kernel<NM_, U_, A_><<<grid, SLABP_THREADS, 0, stream>>>(m, v, colour, sweep, pa, np, pb, partial)
    if (!update) PP_SLABP(3, false, true);
    else if (nm == 3) { if (acc) PP_SLABP(3, true, true); else PP_SLABP(3, true, false); }
    else { if (acc) PP_SLABP(7, true, true); else PP_SLABP(7, true, false); }
#undef PP_SLABP
    s->launches++;
    return PP_OK;
}

// bit-packed sweeps; want_energy: the colour-1 pass of the last sweep also counts bonds and spins (n_sweeps = 0: a counting
// pass over the current state), and the totals become energies / magnetisations
static pp_status slabp_sweeps(pp_sim *s, const ModelView &m, cudaStream_t stream, uint32_t sweep_index, int n_sweeps, int sweep_mode,
                              bool want_energy, bool want_mags) {
    SlabState *sl = s->slab;
    const int nm = sweep_mode == PP_SWEEP_GIBBS ? sl->nm_gibbs : sl->nm_metro;
    if (n_sweeps > 0 && !(sweep_mode == PP_SWEEP_GIBBS ? sl->mono_gibbs : sl->mono_metro))
        return fail(PP_ERR_UNSUPPORTED, "bit-packed slab kernel: the acceptance table does not grow with the number of unsatisfied bonds");
    const uint64_t key = realization_seed(m.seed, (uint64_t)m.sample_offset);
    for (SlabPView &v : sl->pparts) { v.k0 = (uint32_t)key; v.k1 = (uint32_t)(key >> 32); }
    const bool multi = sl->pparts.size() > 1 || sl->comm;
    pp_status st = PP_OK;
    if (want_energy) CUDA_TRY(cudaMemsetAsync(sl->d_partial, 0, sizeof(unsigned long long) * 2 * (size_t)m.S, stream));
    if (n_sweeps == 0 && want_energy) {
        for (const SlabPView &v : sl->pparts)
            if ((st = slabp_launch(s, m, v, stream, 1, 0u, 1, sl->P, 0, 3, false, true)) != PP_OK) return st;
        return slab_finish_energy(s, m, stream, want_mags);
    }
    for (int sw = 0; sw < n_sweeps; sw++)
        for (int colour = 0; colour < 2; colour++) {
            const bool acc = want_energy && sw == n_sweeps - 1 && colour == 1;
            const uint32_t sweep = sweep_index + (uint32_t)sw;
            if (!multi) {  // one slab: every plane in one launch, then the periodic images
                if ((st = slabp_launch(s, m, sl->pparts[0], stream, colour, sweep, 1, sl->P, 0, nm, true, acc)) != PP_OK) return st;
                if ((st = slab_exchange(s, stream, colour)) != PP_OK) return st;
                continue;
            }
            // boundary planes + their transfer on the comm stream, interior planes on the main stream, side by side
            CUDA_TRY(cudaEventRecord(sl->ev_main, stream));
            CUDA_TRY(cudaStreamWaitEvent(sl->comm_stream, sl->ev_main, 0));
            for (const SlabPView &v : sl->pparts)
                if ((st = slabp_launch(s, m, v, sl->comm_stream, colour, sweep, 1, 1, sl->P, nm, true, acc)) != PP_OK) return st;
            if ((st = slab_exchange(s, sl->comm_stream, colour)) != PP_OK) return st;
            CUDA_TRY(cudaEventRecord(sl->ev_halo, sl->comm_stream));
            for (const SlabPView &v : sl->pparts)
                if ((st = slabp_launch(s, m, v, stream, colour, sweep, 2, sl->P - 2, 0, nm, true, acc)) != PP_OK) return st;
            CUDA_TRY(cudaStreamWaitEvent(stream, sl->ev_halo, 0));
        }
    CUDA_TRY(cudaGetLastError());
    return want_energy ? slab_finish_energy(s, m, stream, want_mags) : PP_OK;
}

static pp_status slab_sweeps(pp_sim *s, const ModelView &m, cudaStream_t stream, uint32_t sweep_index, int n_sweeps) {
    SlabState *sl = s->slab;
    const SlabView &v0 = sl->parts[0];
    const unsigned tiles = (unsigned)(((v0.L2 / 8 + v0.tile_x - 1) / v0.tile_x) * ((v0.L1 + 256 / v0.tile_x - 1) / (256 / v0.tile_x)));
    const uint64_t key = realization_seed(m.seed, (uint64_t)m.sample_offset);
    for (SlabView &v : sl->parts) { v.k0 = (uint32_t)key; v.k1 = (uint32_t)(key >> 32); }
    for (int sw = 0; sw < n_sweeps; sw++)
        for (int colour = 0; colour < 2; colour++) {
            for (const SlabView &v : sl->parts) {  // boundary planes (one launch) first: they are what the neighbours wait for
                slab_sweep_kernel<<<dim3(tiles, 2, (unsigned)m.S), 256, 0, stream>>>(m, v, colour, sweep_index + (uint32_t)sw, 1, 1, sl->P);
                s->launches++;
            }
            CUDA_TRY(cudaEventRecord(sl->ev_boundary, stream));
            CUDA_TRY(cudaStreamWaitEvent(sl->comm_stream, sl->ev_boundary, 0));
            pp_status st = slab_exchange(s, sl->comm_stream);
            if (st != PP_OK) return st;
            CUDA_TRY(cudaEventRecord(sl->ev_halo, sl->comm_stream));
            if (sl->P > 2)
                for (const SlabView &v : sl->parts) {  // interior planes overlap the halo transfer
                    slab_sweep_kernel<<<dim3(tiles, (unsigned)((sl->P - 2 + SLAB_PR - 1) / SLAB_PR), (unsigned)m.S), 256, 0, stream>>>(
                        m, v, colour, sweep_index + (uint32_t)sw, 2, sl->P - 2, 0);
                    s->launches++;
                }
            CUDA_TRY(cudaStreamWaitEvent(stream, sl->ev_halo, 0));
        }
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

static pp_status slab_energy(pp_sim *s, const ModelView &m, cudaStream_t stream, bool want_mags) {
    SlabState *sl = s->slab;
    if (sl->packed) return slabp_sweeps(s, m, stream, 0u, 0, PP_SWEEP_METROPOLIS, true, want_mags);
    CUDA_TRY(cudaMemsetAsync(sl->d_partial, 0, sizeof(unsigned long long) * 2 * (size_t)m.S, stream));
    const unsigned bx = blocks_for(sl->parts[0].chunks_per_plane, 256);
    for (const SlabView &v : sl->parts) {
        slab_energy_kernel<<<dim3(bx, (unsigned)sl->P, (unsigned)m.S), 256, 0, stream>>>(v, sl->d_partial);
        s->launches++;
    }
    return slab_finish_energy(s, m, stream, want_mags);
}

// ---- ferromagnets with one bit per spin (pp_kernels_prows.cuh) --------------------------------------------------------------
// packed words <-> the int8 view; dir 0: pack (int8 -> words), 1: unpack
static pp_status prows_sync(pp_sim *s, cudaStream_t stream, int dir) {
    const ModelView &m = s->mv;
    prows_convert_kernel<<<dim3((unsigned)(m.D * m.S), blocks_for(s->rv.n_rows * s->pv.W, PROWS_THREADS)), PROWS_THREADS, 0, stream>>>(m, s->rv, s->pv, dir);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// rec_from >= 0 (cluster mode, s->prows_cluster): the sweeps sw >= rec_from of the batch are recorded inside the kernel (energies,
// magnetisations, pair dots, fold) by thread-block clusters of R CTAs, one per temperature slot
static pp_status launch_prows(pp_sim *s, Ctx &c, const ModelView &m_in, int sweep_mode, uint32_t sweep_index, int n_sweeps, bool want_energy,
                              bool want_mags, int rec_from = -1) {
    ModelView m = m_in;
    m.lut = sweep_mode == PP_SWEEP_GIBBS ? s->d_lut_gibbs : s->d_lut_metro;  // (the energy-only call comes without a table)
    RowsView v = s->rv;
    v.keys = s->d_keys + (c.m.sample_offset - s->mv.sample_offset);
    const int nm = s->prows_nm[sweep_mode == PP_SWEEP_GIBBS ? 1 : 0];
    const bool bs = m.z == 3 && nm == 7 && s->prows_bs[sweep_mode == PP_SWEEP_GIBBS ? 1 : 0];
    PRowsCluster cl{};
    cl.on = rec_from >= 0 ? 1 : 0;
    cl.rec_from = rec_from >= 0 ? rec_from : 0x7FFFFFFF;
    cl.st = c.st;
    cl.dot_spin = c.dot_spin;
    cl.dot_link = c.dot_link;
    size_t smem = sizeof(uint32_t) * (size_t)s->pv.sys_words * (cl.on ? 2 : 1);
    const size_t tab_bytes = (size_t)s->rv.n_rows * (2 * (size_t)m.z + 2) * sizeof(uint32_t) + (((size_t)s->rv.n_rows + 15) & ~size_t(15));
    cl.stage_tables = tab_bytes <= 64 * 1024 && smem + tab_bytes <= 200 * 1024 ? 1 : 0;
    if (cl.stage_tables) smem += tab_bytes;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(m.D * m.S));
    cfg.blockDim = dim3(PROWS_SWEEP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)m.R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl.on ? 1 : 0;
    const int we = want_energy ? 1 : 0, wm = want_mags ? 1 : 0;
#define PP_PROWS(Z_, NM_, BS_)                                                                                                         \
    do {                                                                                                                               \
        static bool configured = false;                                                                                                \
        if (!configured) {                                                                                                             \
            CUDA_TRY(cudaFuncSetAttribute(prows_sweep_kernel<Z_, NM_, BS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));  \
            configured = true;                                                                                                         \
        }                                                                                                                              \
        CUDA_TRY(cudaLaunchKernelEx(&cfg, prows_sweep_kernel<Z_, NM_, BS_>, m, v, s->pv, sweep_index, n_sweeps, we, wm, cl));           \
    } while (0)
    if (m.z == 2) { if (nm == 2) PP_PROWS(2, 2, false); else PP_PROWS(2, 5, false); }
    else if (m.z == 3) { if (nm == 3) PP_PROWS(3, 3, false); else if (bs) PP_PROWS(3, 7, true); else PP_PROWS(3, 7, false); }
    else { if (nm == 4) PP_PROWS(4, 4, false); else PP_PROWS(4, 9, false); }
#undef PP_PROWS
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// ---- fp32 couplings, the same site of 32 systems in one word (pp_kernels_swords.cuh) -----------------------------------------
static pp_status swords_sync(pp_sim *s, cudaStream_t stream, int dir) {
    const ModelView &m = s->mv;
    swords_convert_kernel<<<dim3((unsigned)(m.D * s->swv.SW), blocks_for(m.N, 256)), 256, 0, stream>>>(m, s->swv, dir);
    s->launches++;
    if (dir == 0) s->sw_tbits_valid = false;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// the packed state <-> the int8 view that the API and the int8-only kernels (cluster moves) see; dir 0: pack, 1: unpack
static pp_status view_sync(pp_sim *s, cudaStream_t stream, int dir) {
    if (s->prows) return prows_sync(s, stream, dir);
    if (s->swords) return swords_sync(s, stream, dir);
    return PP_OK;
}

static SWordsView swords_view(pp_sim *s, const Ctx &c) {
    SWordsView sv = s->swv;
    const int64_t d0 = c.m.sample_offset - s->mv.sample_offset;
    const size_t nq = (size_t)(c.m.N / 32);
    sv.words += (size_t)d0 * sv.SW * c.m.N;
    sv.tbits += (size_t)d0 * c.m.S * nq;
    sv.acc_e += (size_t)d0 * sv.SW * 32;
    sv.acc_m += (size_t)d0 * sv.SW * 32;
    sv.arrive_e += (size_t)d0 * sv.SW;
    sv.arrive_m += (size_t)d0 * sv.SW;
    sv.lane_t += (size_t)d0 * sv.SW * 32;
    return sv;
}

static pp_status swords_transpose(pp_sim *s, Ctx &c, bool want_mags) {
    const ModelView &m = c.m;
    const SWordsView sv = swords_view(s, c);
    swords_transpose_kernel<<<dim3((unsigned)(m.D * sv.SW), blocks_for(m.N / 32, SW_THREADS)), SW_THREADS, 0, c.stream>>>(m, sv, want_mags ? 1 : 0);
    s->launches++;
    s->sw_tbits_valid = true;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// n_sweeps colour-pass pairs; the last colour pass also delivers the energies when they are wanted (n_sweeps == 0: the bond sums of
// the last colour's sites without an update); magnetisations come with the transposed view
static pp_status launch_swords(pp_sim *s, Ctx &c, const ModelView &m, int sweep_mode, uint32_t sweep_index, int n_sweeps, int exact_log,
                               bool want_energy, bool want_mags) {
    RowsView v = s->rv;
    v.keys = s->d_keys + (c.m.sample_offset - s->mv.sample_offset);
    const SWordsView sv = swords_view(s, c);
    // sites per thread: SW_SPT when the batch is large, fewer while the grid would not fill the GPU a few times over
    int spt = SW_SPT;
    while (spt > 1 && (int64_t)m.D * sv.SW * (int64_t)blocks_for(m.N / 2, SW_THREADS * spt) < 4 * 148 * 8) spt /= 2;
    const dim3 grid((unsigned)(m.D * sv.SW), blocks_for(m.N / 2, SW_THREADS * spt));
    const bool gibbs = sweep_mode == PP_SWEEP_GIBBS;
    if (n_sweeps > 0) {  // system_ids only change between launch sequences (parallel tempering, set_system_ids)
        swords_lane_temps_kernel<<<blocks_for(m.D * m.S, 256), 256, 0, c.stream>>>(m, sv);
        s->launches++;
    }
#define PP_SW5(Z_, G_, U_, E_, X_) swords_sweep_kernel<Z_, G_, U_, E_, X_><<<grid, SW_THREADS, 0, c.stream>>>(m, v, sv, col, sweep_index + (uint32_t)sw, spt)
#define PP_SW4(Z_, G_, U_, E_) do { if (exact_log) PP_SW5(Z_, G_, U_, E_, true); else PP_SW5(Z_, G_, U_, E_, false); } while (0)
#define PP_SW3(Z_, G_) do { if (!update) PP_SW5(Z_, false, false, true, false); else if (eacc) PP_SW4(Z_, G_, true, true); else PP_SW4(Z_, G_, true, false); } while (0)
#define PP_SW2(Z_) do { if (gibbs) PP_SW3(Z_, true); else PP_SW3(Z_, false); } while (0)
    for (int sw = 0; sw < std::max(n_sweeps, 1); sw++)
        for (int col = 0; col < 2; col++) {
            const bool update = n_sweeps > 0;
            const bool eacc = want_energy && sw == std::max(n_sweeps, 1) - 1 && col == 1;
            if (!update && !eacc) continue;
            if (m.z == 2) PP_SW2(2); else PP_SW2(3);
            s->launches++;
        }
#undef PP_SW5
#undef PP_SW4
#undef PP_SW3
#undef PP_SW2
    if (n_sweeps > 0) s->sw_tbits_valid = false;
    CUDA_TRY(cudaGetLastError());
    if (want_mags) return swords_transpose(s, c, true);
    return PP_OK;
}

// want_overlap / want_fold: the caller wants the replica-pair dots of the post-sweep state / the recorded-sweep fold;
// *fused is set when the sweep kernel did both itself (msc3d epilogue), otherwise the caller launches
// launch_overlap() and fold_kernel.
static pp_status launch_sweeps(pp_sim *s, Ctx &c, int sweep_mode, uint32_t sweep_index, int n_sweeps, int exact_log,
                               bool want_energy, bool want_mags, bool want_overlap = false, bool want_fold = false,
                               bool *fused = nullptr, int rec_from = -1) {
    ModelView m = c.m;
    m.lut = sweep_mode == PP_SWEEP_GIBBS ? s->d_lut_gibbs : s->d_lut_metro;
    if (fused) *fused = false;
    if (s->layout == PP_LAYOUT_SLAB) {
        if (n_sweeps > 0) prof_mark(s, c.stream);
        // bit-packed storage: bond / spin counts come out of the last colour pass (no second pass over the lattice)
        pp_status st = s->slab->packed ? slabp_sweeps(s, m, c.stream, sweep_index, n_sweeps, sweep_mode, want_energy, want_mags)
                                       : slab_sweeps(s, m, c.stream, sweep_index, n_sweeps);
        if (st != PP_OK) return st;
        if (n_sweeps > 0) prof_mark(s, c.stream);
        return want_energy && !s->slab->packed ? slab_energy(s, m, c.stream, want_mags) : PP_OK;
    }
    if (s->layout == PP_LAYOUT_MSC) {
        const bool timed = n_sweeps > 0;
        if (timed) prof_mark(s, c.stream);
        if (s->msc3d) {
            const bool ov = want_overlap && m.P > 0;
            pp_status st = launch_msc3d(s, c, m, sweep_mode, sweep_index, n_sweeps, want_energy, want_mags, ov, want_fold);
            if (st != PP_OK) return st;
            if (fused) *fused = true;
            if (timed) prof_mark(s, c.stream);
            return PP_OK;
        }
        const size_t smem = sizeof(uint32_t) * (size_t)m.N;
        const unsigned grid = (unsigned)(c.G * m.S);
        if (smem <= 200 * 1024) {
            CUDA_TRY(cudaFuncSetAttribute(msc_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            msc_sweep_kernel<true><<<grid, MSC_BLOCK, smem, c.stream>>>(m, sweep_index, n_sweeps, want_energy, want_mags,
                                                                       m.sample_offset / 32);
        } else {
            msc_sweep_kernel<false><<<grid, MSC_BLOCK, 0, c.stream>>>(m, sweep_index, n_sweeps, want_energy, want_mags,
                                                                     m.sample_offset / 32);
        }
        s->launches++;
        if (timed) prof_mark(s, c.stream);
        CUDA_TRY(cudaGetLastError());
        return PP_OK;
    }
    if (n_sweeps > 0) prof_mark(s, c.stream);
    if (s->prows) {  // every sweep of the batch in one launch, the system in shared memory; energies out of the same launch
        const bool cluster = s->prows_cluster && rec_from >= 0 && n_sweeps > 0 && want_fold;
        pp_status st = launch_prows(s, c, m, sweep_mode, sweep_index, n_sweeps, want_energy, want_mags, cluster ? rec_from : -1);
        if (cluster && fused) *fused = true;  // energies, magnetisations, pair dots and the fold of the recorded sweeps are done
        if (n_sweeps > 0) prof_mark(s, c.stream);
        return st;
    }
    if (s->swords) {
        pp_status st = launch_swords(s, c, m, sweep_mode, sweep_index, n_sweeps, exact_log, want_energy, want_mags);
        if (n_sweeps > 0) prof_mark(s, c.stream);
        return st;
    }
    if (s->rows) {
        RowsView v = s->rv;
        v.keys = s->d_keys + (c.m.sample_offset - s->mv.sample_offset);
        const size_t smem = m.coupling_class == COUP_F32 ? 0 : sizeof(uint32_t) * (size_t)m.T * (4 * m.z + 1);
        const int ns = m.coupling_class == COUP_FERRO ? rows_ns<COUP_FERRO>() : ROWS_NS;
        const int sblocks = (m.S + ns - 1) / ns;
        // two-colour lattices: the last colour pass of the last sweep also delivers the energies (+ magnetisation sums)
        const bool esw = s->rows_esw && want_energy && n_sweeps > 0 && m.n_colours == 2;
        bool energy_done = false;
        for (int sw = 0; sw < n_sweeps; sw++)
            for (int col = 0; col < m.n_colours; col++) {
                const int cls = col % v.m_half;
                const uint32_t nseg = (s->rows_class_start[(size_t)cls + 1] - s->rows_class_start[(size_t)cls]) * (uint32_t)v.kpr;
                dim3 grid((unsigned)(m.D * sblocks), blocks_for(nseg, 128));
                const bool eacc = esw && sw == n_sweeps - 1 && col == m.n_colours - 1;
#define PP_ROWS3(C_, Z_, G_, E_) rows_sweep_kernel<C_, Z_, G_, E_><<<grid, 128, smem, c.stream>>>(m, v, col, sweep_index + sw, exact_log, s->rows_escale, want_mags ? 1 : 0, s->d_rows_acc, s->d_rows_arrive)
#define PP_ROWS2(C_, Z_, G_) do { if (eacc) PP_ROWS3(C_, Z_, G_, true); else PP_ROWS3(C_, Z_, G_, false); } while (0)
#define PP_ROWS(C_, Z_) do { if (sweep_mode == PP_SWEEP_GIBBS) PP_ROWS2(C_, Z_, true); else PP_ROWS2(C_, Z_, false); } while (0)
                if (m.coupling_class == COUP_FERRO) { if (m.z == 2) PP_ROWS(COUP_FERRO, 2); else if (m.z == 3) PP_ROWS(COUP_FERRO, 3); else PP_ROWS(COUP_FERRO, 0); }
                else if (m.coupling_class == COUP_UNIT) { if (m.z == 2) PP_ROWS(COUP_UNIT, 2); else if (m.z == 3) PP_ROWS(COUP_UNIT, 3); else PP_ROWS(COUP_UNIT, 0); }
                else { if (m.z == 2) PP_ROWS(COUP_F32, 2); else if (m.z == 3) PP_ROWS(COUP_F32, 3); else PP_ROWS(COUP_F32, 0); }
#undef PP_ROWS3
#undef PP_ROWS2
#undef PP_ROWS
                s->launches++;
                if (eacc) energy_done = true;
            }
        if (n_sweeps > 0) prof_mark(s, c.stream);
        CUDA_TRY(cudaGetLastError());
        if (want_energy && !energy_done) return launch_energy(s, c, want_mags);
        return PP_OK;
    }
    for (int sw = 0; sw < n_sweeps; sw++) {
        for (int col = 0; col < m.n_colours; col++) {
            const uint32_t nsite = s->plan.colour_start[col + 1] - s->plan.colour_start[col];
            if (nsite == 0) continue;
            dim3 grid((unsigned)(m.D * m.S), blocks_for((nsite + 3) / 4, SWEEP_BLOCK));
            switch (m.coupling_class) {
                case COUP_FERRO:
                    sweep_colour_int8_kernel<COUP_FERRO><<<grid, SWEEP_BLOCK, 0, c.stream>>>(m, col, sweep_index + sw, sweep_mode, exact_log);
                    break;
                case COUP_UNIT:
                    sweep_colour_int8_kernel<COUP_UNIT><<<grid, SWEEP_BLOCK, 0, c.stream>>>(m, col, sweep_index + sw, sweep_mode, exact_log);
                    break;
                default:
                    sweep_colour_int8_kernel<COUP_F32><<<grid, SWEEP_BLOCK, 0, c.stream>>>(m, col, sweep_index + sw, sweep_mode, exact_log);
            }
            s->launches++;
        }
    }
    if (n_sweeps > 0) prof_mark(s, c.stream);
    CUDA_TRY(cudaGetLastError());
    if (want_energy) return launch_energy(s, c, want_mags);
    return PP_OK;
}

static pp_status launch_energy(pp_sim *s, Ctx &c, bool want_mags) {
    const ModelView &m = c.m;
    if (s->layout == PP_LAYOUT_MSC) return launch_sweeps(s, c, PP_SWEEP_METROPOLIS, 0, 0, 0, true, want_mags);
    if (s->layout == PP_LAYOUT_SLAB) return slab_energy(s, m, c.stream, want_mags);
    const unsigned grid = (unsigned)(m.D * m.S);
    if (s->prows) return launch_prows(s, c, m, PP_SWEEP_METROPOLIS, 0u, 0, true, want_mags);
    if (s->swords) return launch_swords(s, c, m, PP_SWEEP_METROPOLIS, 0u, 0, 0, true, want_mags);
    if (s->rows) {
        const dim3 g2(grid, (unsigned)(m.coupling_class == COUP_F32 ? 1 : s->rows_nb));
#define PP_RE(C_, Z_) rows_energy_kernel<C_, Z_><<<g2, 256, 0, c.stream>>>(m, s->rv, want_mags, s->d_rows_acc, s->d_rows_arrive)
#define PP_REZ(C_) do { if (m.z == 2) PP_RE(C_, 2); else if (m.z == 3) PP_RE(C_, 3); else PP_RE(C_, 0); } while (0)
        switch (m.coupling_class) {
            case COUP_FERRO: PP_REZ(COUP_FERRO); break;
            case COUP_UNIT: PP_REZ(COUP_UNIT); break;
            default: PP_REZ(COUP_F32);
        }
#undef PP_REZ
#undef PP_RE
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return PP_OK;
    }
    switch (m.coupling_class) {
        case COUP_FERRO: energy_mag_int8_kernel<COUP_FERRO><<<grid, 256, 0, c.stream>>>(m, want_mags); break;
        case COUP_UNIT: energy_mag_int8_kernel<COUP_UNIT><<<grid, 256, 0, c.stream>>>(m, want_mags); break;
        default: energy_mag_int8_kernel<COUP_F32><<<grid, 256, 0, c.stream>>>(m, want_mags);
    }
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// fold_too / *folded: the packed-row overlap kernel also runs the recorded-sweep fold (also when there are no replica pairs)
static pp_status launch_overlap(pp_sim *s, Ctx &c, bool fold_too = false, bool *folded = nullptr) {
    ModelView m = c.m;
    if (folded) *folded = false;
    if (s->prows && s->layout == PP_LAYOUT_INT8 && (m.P > 0 || fold_too)) {
        const unsigned grid = (unsigned)(m.D * m.T);
        const int wf = fold_too ? 1 : 0;
        if (m.z == 2) prows_overlap_kernel<2><<<grid, PROWS_THREADS, 0, c.stream>>>(m, s->rv, s->pv, c.st, c.dot_spin, c.dot_link, wf);
        else if (m.z == 3) prows_overlap_kernel<3><<<grid, PROWS_THREADS, 0, c.stream>>>(m, s->rv, s->pv, c.st, c.dot_spin, c.dot_link, wf);
        else prows_overlap_kernel<4><<<grid, PROWS_THREADS, 0, c.stream>>>(m, s->rv, s->pv, c.st, c.dot_spin, c.dot_link, wf);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        if (folded) *folded = fold_too;
        return PP_OK;
    }
    if (m.P == 0) return PP_OK;
    if (s->swords) {
        if (!s->sw_tbits_valid) {
            pp_status st = swords_transpose(s, c, false);
            if (st != PP_OK) return st;
        }
        const SWordsView sv = swords_view(s, c);
        const size_t per_warp = (size_t)(m.N / 32) * sizeof(uint32_t);
        const int wpc = (int)std::max<size_t>(1, std::min<size_t>(8, (160 * 1024) / per_warp));  // warps (= pairs) per CTA
        const unsigned grid = blocks_for(m.D * m.P * m.T, wpc);
        const size_t smem = per_warp * (size_t)wpc;
        if (smem > 48 * 1024) {
            CUDA_TRY(cudaFuncSetAttribute(swords_overlap_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(swords_overlap_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        if (m.z == 2) swords_overlap_kernel<2><<<grid, 32 * wpc, smem, c.stream>>>(m, s->rv, sv, c.dot_spin, c.dot_link);
        else swords_overlap_kernel<3><<<grid, 32 * wpc, smem, c.stream>>>(m, s->rv, sv, c.dot_spin, c.dot_link);
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        return PP_OK;
    }
    if (s->layout == PP_LAYOUT_MSC && s->msc3d) {
        m.lut = s->d_lut_metro;
        return launch_msc3d(s, c, m, PP_SWEEP_METROPOLIS, 0, 0, false, false, true, false);
    }
    if (s->layout == PP_LAYOUT_MSC)
        msc_overlap_kernel<<<(unsigned)(c.G * m.P * m.T), MSC_BLOCK, 0, c.stream>>>(m, c.dot_spin, c.dot_link);
    else if (s->rows)
        rows_overlap_kernel<<<dim3((unsigned)(m.D * m.P * m.T), (unsigned)s->rows_nb), 256, 0, c.stream>>>(m, s->rv, c.dot_spin, c.dot_link,
                                                                                                   s->d_rows_acc, s->d_rows_arrive);
    else
        overlap_dots_int8_kernel<<<(unsigned)(m.D * m.P * m.T), 256, 0, c.stream>>>(m, c.dot_spin, c.dot_link);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// one parallel-tempering event (mod.rs:748-796); first_parity is the caller's PtState.next_parity
// exchange the lanes of a pending parallel-tempering event in place (the msc3d sweep kernel normally does this while
// staging its input; this is the path for everything else that looks at the words)
static pp_status flush_swaps(pp_sim *s, Ctx &c) {
    if (!c.swap_pending) return PP_OK;
    const ModelView &m = c.m;
    dim3 grid((unsigned)(c.G * m.R), blocks_for(m.N, 256));
    msc_apply_swaps_kernel<<<grid, 256, 0, c.stream>>>(m, c.pt.swap_mask, c.pend_schedule, c.pend_parity);
    s->launches++;
    c.swap_pending = false;
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// defer = the caller guarantees that the next kernel touching the words is an msc3d sweep (or flush_swaps)
static pp_status launch_pt(pp_sim *s, Ctx &c, int schedule, uint32_t pt_event, int first_parity, bool defer = false) {
    const ModelView &m = c.m;
    if (m.T < 2) return PP_OK;
    if (s->layout == PP_LAYOUT_MSC) {
        pp_status st = flush_swaps(s, c);
        if (st != PP_OK) return st;
        pt_exchange_msc_kernel<<<blocks_for(c.G * m.R * 32, 128), 128, 0, c.stream>>>(m, c.pt, schedule, first_parity, pt_event);
        s->launches++;
        c.swap_pending = true;
        c.pend_schedule = schedule;
        c.pend_parity = first_parity;
        if (!(defer && s->msc3d && s->defer_swaps)) {
            st = flush_swaps(s, c);
            if (st != PP_OK) return st;
        }
    } else {
        pt_exchange_kernel<<<blocks_for(m.D * m.R, 128), 128, 0, c.stream>>>(m, c.pt, schedule, first_parity, pt_event);
        s->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return PP_OK;
}

// system-split handle: every rank contributes what it produced for its own systems (contiguous blocks, in place)
static pp_status sys_allgather(pp_sim *s, cudaStream_t stream, bool energies, bool mags, bool spins) {
    if (!s->sys_comm) return PP_OK;
    NcclApi &nc = nccl_api();
    const ModelView &m = s->mv;
    const size_t per = (size_t)(m.S / s->sys_ranks);
    if (energies)
        NCCL_TRY(nc.AllGather((const char *)(s->d_energies + m.sys_lo), s->d_energies, per * sizeof(float), ncclUint8, s->sys_comm, stream));
    if (mags)
        NCCL_TRY(nc.AllGather((const char *)(s->d_mags + m.sys_lo), s->d_mags, per * sizeof(long long), ncclUint8, s->sys_comm, stream));
    if (spins && s->prows)
        NCCL_TRY(nc.AllGather((const char *)(s->pv.words + (size_t)m.sys_lo * s->pv.sys_words), s->pv.words,
                              per * (size_t)s->pv.sys_words * sizeof(uint32_t), ncclUint8, s->sys_comm, stream));
    else if (spins)
        NCCL_TRY(nc.AllGather((const char *)(s->d_spins + (size_t)m.sys_lo * m.N), s->d_spins, per * (size_t)m.N, ncclUint8, s->sys_comm, stream));
    s->launches += (energies ? 1 : 0) + (mags ? 1 : 0) + (spins ? 1 : 0);
    return PP_OK;
}

// Realization::new / reset (realization.rs:166-207, 213-246) on the device
static pp_status do_reset(pp_sim *s, uint64_t seed) {
    CUDA_TRY(cudaSetDevice(s->device));
    s->mv.seed = seed;
    s->sweep_counter = 0;
    s->pt_event_counter = 0;
    s->next_parity = 0;
    ModelView m = s->mv;
    const int64_t DS = m.D * m.S;
    if (s->rows) {
        std::vector<uint64_t> keys((size_t)m.D);
        for (int64_t d = 0; d < m.D; d++) keys[(size_t)d] = realization_seed(seed, (uint64_t)(m.sample_offset + d));
        CUDA_TRY(cudaMemcpyAsync(s->d_keys, keys.data(), sizeof(uint64_t) * keys.size(), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    iota_sid_kernel<<<blocks_for(DS, 256), 256, 0, s->stream>>>(s->d_sid, DS, m.S);
    if (s->layout == PP_LAYOUT_SLAB) {
        const uint64_t key = realization_seed(m.seed, (uint64_t)m.sample_offset);
        for (SlabView &v : s->slab->parts) {
            v.k0 = (uint32_t)key;
            v.k1 = (uint32_t)(key >> 32);
            slab_init_kernel<<<dim3(blocks_for(v.chunks_per_plane, 256), (unsigned)v.P, (unsigned)m.S), 256, 0, s->stream>>>(m, v);
        }
        for (SlabPView &v : s->slab->pparts) {
            v.k0 = (uint32_t)key;
            v.k1 = (uint32_t)(key >> 32);
            slabp_init_kernel<<<dim3(blocks_for(v.half, SLABP_THREADS), (unsigned)v.P, (unsigned)m.S), SLABP_THREADS, 0, s->stream>>>(m, v);
        }
        pp_status stx = slab_exchange(s, s->stream);
        if (stx != PP_OK) return stx;
    } else if (s->layout == PP_LAYOUT_MSC) {
        dim3 grid((unsigned)(s->G * m.S), blocks_for((m.N + 3) / 4, 128));
        msc_init_kernel<<<grid, 128, 0, s->stream>>>(m);
    } else {
        dim3 grid((unsigned)DS, blocks_for((m.N + 3) / 4, 128));
        init_spins_int8_kernel<<<grid, 128, 0, s->stream>>>(m);
        if (s->prows || s->swords) {  // the same site-indexed INIT draws as every layout, then one bit per spin
            pp_status stp = view_sync(s, s->stream, 0);
            if (stp != PP_OK) return stp;
        }
    }
    CUDA_TRY(cudaGetLastError());
    Ctx wc = whole_ctx(s);
    pp_status stt = launch_energy(s, wc, false);
    if (stt != PP_OK) return stt;
    if (m.T > 1) {
        CUDA_TRY(cudaMemsetAsync(s->pt.edge_attempts, 0, sizeof(uint64_t) * (size_t)(m.D * (m.T - 1)), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->pt.edge_acceptances, 0, sizeof(uint64_t) * (size_t)(m.D * (m.T - 1)), s->stream));
    }
    CUDA_TRY(cudaMemsetAsync(s->pt.round_trips, 0, sizeof(uint64_t) * (size_t)DS, s->stream));
    CUDA_TRY(cudaMemsetAsync(s->pt.trip_state, 0, (size_t)DS, s->stream));
    pt_mark_hot_kernel<<<blocks_for(m.D * m.R, 128), 128, 0, s->stream>>>(m, s->pt);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return PP_OK;
}

extern "C" pp_status pp_create(const pp_model_desc *desc, pp_sim **out) {
    if (!desc || !out) return fail(PP_ERR_INVALID, "desc/out is NULL");
    *out = nullptr;
    if (!desc->shape || !desc->temperatures) return fail(PP_ERR_INVALID, "shape/temperatures is NULL");
    if (desc->n_temps < 1) return fail(PP_ERR_INVALID, "n_temps must be >= 1");
    if (desc->n_replicas < 1) return fail(PP_ERR_INVALID, "n_replicas must be >= 1");
    if (desc->n_disorder < 1) return fail(PP_ERR_INVALID, "n_disorder must be >= 1");
    if (desc->coupling_kind == PP_COUPLINGS_ARRAY && !desc->couplings) return fail(PP_ERR_INVALID, "couplings is NULL");
    if ((int64_t)desc->n_temps * desc->n_replicas > (1 << 20)) return fail(PP_ERR_INVALID, "too many systems per realization");

    // slab layout (pp_slab.cuh): explicit, or chosen for one large ferromagnet whose neighbour tables would not fit
    const bool slab_shape = desc->n_dims == 3 && (desc->n_offsets <= 0 || !desc->offsets) && desc->shape[0] % 2 == 0 &&
                            desc->shape[1] % 2 == 0 && desc->shape[2] % 8 == 0;
    bool want_slab = desc->layout == PP_LAYOUT_SLAB;
    if (desc->layout == PP_LAYOUT_AUTO && slab_shape && desc->coupling_kind == PP_COUPLINGS_FERRO && desc->n_disorder == 1 &&
        desc->n_replicas == 1 && desc->shape[0] * desc->shape[1] * desc->shape[2] >= (int64_t(1) << 24))
        want_slab = true;
    const int slab_ranks = std::max(1, (int)desc->slab_ranks);
    if (want_slab) {
        if (!slab_shape) return fail(PP_ERR_UNSUPPORTED, "slab layout needs a 3-D hypercubic lattice with even extents and shape[2] % 8 == 0");
        if (desc->coupling_kind != PP_COUPLINGS_FERRO) return fail(PP_ERR_UNSUPPORTED, "slab layout needs couplings='ferro' (never materialised)");
        if (desc->n_disorder != 1 || desc->n_replicas != 1) return fail(PP_ERR_UNSUPPORTED, "slab layout holds one realization with one replica per temperature");
        if (desc->shape[0] % (2 * slab_ranks) != 0) return fail(PP_ERR_INVALID, "slab layout needs shape[0] to be a multiple of 2 * slab_ranks");
        if (desc->slab_rank >= slab_ranks) return fail(PP_ERR_INVALID, "slab_rank out of range");
    } else if (desc->slab_ranks > 1) {
        return fail(PP_ERR_INVALID, "slab_ranks > 1 needs layout = PP_LAYOUT_SLAB");
    }
    const int sys_ranks = std::max(1, (int)desc->system_ranks);
    if (sys_ranks > 1) {
        if (want_slab) return fail(PP_ERR_INVALID, "system_ranks > 1 and the slab layout are two different decompositions");
        if (desc->n_disorder != 1) return fail(PP_ERR_UNSUPPORTED, "system_ranks > 1 splits the systems of ONE realization (shard realizations with sample_offset instead)");
        if (desc->system_rank < 0 || desc->system_rank >= sys_ranks) return fail(PP_ERR_INVALID, "system_rank out of range");
        if (((int64_t)desc->n_temps * desc->n_replicas) % sys_ranks != 0) return fail(PP_ERR_INVALID, "n_replicas * n_temps must be a multiple of system_ranks");
        if (desc->layout != PP_LAYOUT_AUTO && desc->layout != PP_LAYOUT_INT8) return fail(PP_ERR_UNSUPPORTED, "system_ranks > 1 needs the int8 layout");
    }

    pp_sim *s = new pp_sim();
    std::string err = build_plan(desc->n_dims, desc->shape, desc->n_offsets, desc->offsets, s->plan, !want_slab);
    if (!err.empty()) {
        delete s;
        return fail(PP_ERR_UNSUPPORTED, err);
    }
    s->device = desc->device;
    s->ctor_seed = desc->seed;
    if (const char *e = getenv("PP_STREAMS")) s->n_streams = std::max(1, atoi(e));
    if (const char *e = getenv("PP_CHUNK_MIB")) s->chunk_bytes = (int64_t)std::max(1, atoi(e)) << 20;
    if (const char *e = getenv("PP_MACRO_BATCH")) s->macro_batch = std::max(1, atoi(e));
    if (const char *e = getenv("PP_CHUNK_GROUPS")) s->chunk_groups = std::max(0, atoi(e));
    if (const char *e = getenv("PP_MAX_BATCH")) s->max_batch = std::max(1, atoi(e));
    if (const char *e = getenv("PP_DEFER_SWAPS")) s->defer_swaps = atoi(e) != 0;
    s->temps.assign(desc->temperatures, desc->temperatures + desc->n_temps);
    ModelView &m = s->mv;
    m.N = s->plan.n_spins;
    m.z = s->plan.z;
    m.T = desc->n_temps;
    m.R = desc->n_replicas;
    m.S = m.T * m.R;
    m.P = m.R / 2;
    m.D = desc->n_disorder;
    m.sample_offset = desc->sample_offset;
    m.sys_lo = 0;
    m.sys_hi = m.S;
    m.n_colours = s->plan.n_colours;
    m.seed = desc->seed;

#define CREATE_TRY(expr)                                                          \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            std::string msg = std::string(#expr) + ": " + cudaGetErrorString(_e); \
            free_sim(s);                                                          \
            return fail(_e == cudaErrorMemoryAllocation ? PP_ERR_OOM : PP_ERR_CUDA, msg); \
        }                                                                         \
    } while (0)

    CREATE_TRY(cudaSetDevice(s->device));
    CREATE_TRY(cudaStreamCreate(&s->stream));
    CREATE_TRY(cudaEventCreate(&s->ev0));
    CREATE_TRY(cudaEventCreate(&s->ev1));

    const int64_t N = m.N;
    const int z = m.z;
    // couplings: classify exactly like the reference's lookup gate (sweep.rs:109-118)
    const bool t_ok = temps_eligible(s->temps.data(), m.T);
    const int64_t n_coup = m.D * N * z;
    int flags[3] = {0, 0, 0};
    if (desc->coupling_kind == PP_COUPLINGS_FERRO) {
        if (!t_ok) {
            free_sim(s);
            return fail(PP_ERR_UNSUPPORTED, "ferro couplings with non-positive or non-finite temperatures: pass an explicit coupling array");
        }
        m.coupling_class = COUP_FERRO;
    } else {
        CREATE_TRY(pool_alloc(s, (void **)&s->d_Jf, sizeof(float) * (size_t)n_coup));
        CREATE_TRY(cudaMemcpy(s->d_Jf, desc->couplings, sizeof(float) * (size_t)n_coup, cudaMemcpyHostToDevice));
        int *d_flags = nullptr;
        CREATE_TRY(pool_alloc(s, (void **)&d_flags, sizeof(int) * 3));
        CREATE_TRY(cudaMemset(d_flags, 0, sizeof(int) * 3));
        classify_couplings_kernel<<<1184, 256, 0, s->stream>>>(s->d_Jf, n_coup, d_flags);
        CREATE_TRY(cudaStreamSynchronize(s->stream));
        CREATE_TRY(cudaMemcpy(flags, d_flags, sizeof(int) * 3, cudaMemcpyDeviceToHost));
        pool_free(s, d_flags);
        if (flags[0] || !t_ok) m.coupling_class = COUP_F32;
        else if (!flags[1] && !flags[2]) m.coupling_class = COUP_FERRO;
        else m.coupling_class = COUP_UNIT;
    }

    // layout
    // multispin eligibility: +-1 couplings without zeros, at most 7 forward directions, and the per-thread bit-sliced
    // counters of the table-driven kernels (pp_kernels_msc.cuh, MSC_VC_PLANES planes) must hold ceil(N / 256) * z adds
    const bool msc_ok = m.coupling_class != COUP_F32 && !flags[1] && z <= 7 &&
                        ((N + MSC_BLOCK - 1) / MSC_BLOCK) * (int64_t)z < ((int64_t)1 << MSC_VC_PLANES);
    // One draw serves the 32 realizations of a word (RNG-SPEC).  With distinct disorder that only
    // correlates the noise of the realizations; with IDENTICAL couplings (a ferromagnet) two lanes that meet in a slot stay
    // equal from then on, so the automatic choice never packs identical realizations into one word.
    const bool msc_auto_ok = msc_ok && m.coupling_class == COUP_UNIT;
    if (want_slab) {
        s->layout = PP_LAYOUT_SLAB;
    } else if (desc->layout == PP_LAYOUT_MSC) {
        if (!msc_ok) {
            free_sim(s);
            return fail(PP_ERR_UNSUPPORTED, "multispin layout needs +-1 couplings (no zeros), eligible temperatures, <= 7 forward directions and ceil(N / 256) * z < 2^20");
        }
        if (desc->sample_offset % 32 != 0) {
            free_sim(s);
            return fail(PP_ERR_INVALID, "multispin layout needs sample_offset to be a multiple of 32");
        }
        s->layout = PP_LAYOUT_MSC;
    } else if (desc->layout == PP_LAYOUT_AUTO) {
        s->layout = (msc_auto_ok && m.D >= 32 && desc->sample_offset % 32 == 0) ? PP_LAYOUT_MSC : PP_LAYOUT_INT8;
    } else if (desc->layout == PP_LAYOUT_INT8) {
        s->layout = PP_LAYOUT_INT8;
    } else {
        free_sim(s);
        return fail(PP_ERR_INVALID, "unknown layout");
    }

    // geometry tables; the multispin layout stores words in the plan's compact order (pp_plan.h)
    if (s->layout != PP_LAYOUT_SLAB) {
        const bool storage_space = s->layout == PP_LAYOUT_MSC && s->plan.compact;
        std::vector<uint32_t> nbr_s, order_s;
        if (storage_space) storage_tables(s->plan, nbr_s, order_s);
        const std::vector<uint32_t> &nbr = storage_space ? nbr_s : s->plan.nbr;
        const std::vector<uint32_t> &order = storage_space ? order_s : s->plan.order;
        CREATE_TRY(pool_alloc(s, (void **)&s->d_nbr, sizeof(uint32_t) * (size_t)N * 2 * z));
        CREATE_TRY(cudaMemcpy(s->d_nbr, nbr.data(), sizeof(uint32_t) * (size_t)N * 2 * z, cudaMemcpyHostToDevice));
        CREATE_TRY(pool_alloc(s, (void **)&s->d_order, sizeof(uint32_t) * (size_t)N));
        CREATE_TRY(cudaMemcpy(s->d_order, order.data(), sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice));
        CREATE_TRY(pool_alloc(s, (void **)&s->d_colour_start, sizeof(uint32_t) * (size_t)(m.n_colours + 1)));
        CREATE_TRY(cudaMemcpy(s->d_colour_start, s->plan.colour_start.data(), sizeof(uint32_t) * (size_t)(m.n_colours + 1),
                              cudaMemcpyHostToDevice));
        if (storage_space) {
            CREATE_TRY(pool_alloc(s, (void **)&s->d_perm, sizeof(uint32_t) * (size_t)N));
            CREATE_TRY(cudaMemcpy(s->d_perm, s->plan.perm.data(), sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice));
        }
        m.nbr = s->d_nbr;
        m.order = s->d_order;
        m.colour_start = s->d_colour_start;
        m.perm = s->d_perm;
    }

    if (s->layout == PP_LAYOUT_SLAB) {
        SlabState *sl = s->slab = new SlabState();
        sl->ranks = slab_ranks;
        sl->rank = slab_ranks == 1 ? 0 : desc->slab_rank;
        sl->L0 = (int)desc->shape[0]; sl->L1 = (int)desc->shape[1]; sl->L2 = (int)desc->shape[2];
        sl->P = sl->L0 / slab_ranks;
        sl->plane = (int64_t)sl->L1 * sl->L2;
        const int n_local = sl->rank < 0 ? slab_ranks : 1;
        // one bit per spin whenever the rows split into whole 64-site word pairs (pp_kernels_slabp.cuh); PP_SLAB_BYTES=1 keeps
        // the byte storage (A/B timing)
        sl->packed = sl->L2 % 64 == 0 && !(getenv("PP_SLAB_BYTES") && atoi(getenv("PP_SLAB_BYTES")) != 0);
        for (int i = 0; i < n_local && sl->packed; i++) {
            SlabPView v{};
            v.P = sl->P; v.L1 = sl->L1; v.W = sl->L2 / 64;
            v.half = (int64_t)v.L1 * v.W;
            v.plane = 2 * v.half;
            v.sys_stride = (int64_t)(sl->P + 2) * v.plane;
            v.first_plane = (int64_t)(sl->rank < 0 ? i : sl->rank) * sl->P;
            CREATE_TRY(cudaMalloc((void **)&v.words, sizeof(uint32_t) * (size_t)(m.S * v.sys_stride)));
            sl->pparts.push_back(v);
        }
        for (int i = 0; i < n_local && !sl->packed; i++) {
            SlabView v{};
            v.P = sl->P; v.L1 = sl->L1; v.L2 = sl->L2;
            v.plane = sl->plane;
            v.sys_stride = (int64_t)(sl->P + 2) * sl->plane;
            v.first_plane = (int64_t)(sl->rank < 0 ? i : sl->rank) * sl->P;
            v.chunks_per_plane = sl->plane / 8;
            v.kpr_shift = -1;
            for (int b = 0; b < 30; b++)
                if ((sl->L2 >> 3) == (1 << b)) v.kpr_shift = b;
            v.tile_x_shift = 0;
            while (v.tile_x_shift < 5 && (2 << v.tile_x_shift) <= (sl->L2 >> 3)) v.tile_x_shift++;
            v.tile_x = 1 << v.tile_x_shift;
            uint8_t *buf = nullptr;
            CREATE_TRY(cudaMalloc((void **)&buf, (size_t)(m.S * v.sys_stride)));
            sl->buffers.push_back(buf);
            v.spins = buf;
            sl->parts.push_back(v);
        }
        CREATE_TRY(cudaMalloc((void **)&sl->d_partial, sizeof(unsigned long long) * 2 * (size_t)m.S));
        {   // the boundary planes and their transfer run on a stream of their own, ahead of the interior launches
            int prio_lo = 0, prio_hi = 0;
            CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CREATE_TRY(cudaStreamCreateWithPriority(&sl->comm_stream, cudaStreamNonBlocking, prio_hi));
        }
        CREATE_TRY(cudaEventCreateWithFlags(&sl->ev_boundary, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&sl->ev_halo, cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&sl->ev_main, cudaEventDisableTiming));
        if (slab_ranks > 1 && sl->rank >= 0) {
            pp_status cst = cached_comm(s->device, slab_ranks, sl->rank, desc->nccl_unique_id, &sl->comm);
            if (cst != PP_OK) {
                std::string keep = g_last_error;
                sl->comm = nullptr;
                free_sim(s);
                return fail(cst, keep);
            }
            sl->comm_cached = true;
        }
    } else if (s->layout == PP_LAYOUT_MSC) {
        s->G = (m.D + 31) / 32;
        if (m.coupling_class == COUP_UNIT) {
            CREATE_TRY(pool_alloc(s, (void **)&s->d_Jw, sizeof(uint32_t) * (size_t)(s->G * z * N)));
            msc_pack_couplings_kernel<<<blocks_for(s->G * z * N, 256), 256, 0, s->stream>>>(s->d_Jf, s->d_Jw, m.D, N, z, s->d_perm);
            CREATE_TRY(cudaStreamSynchronize(s->stream));
        } else {  // ferromagnet: all-zero sign words, so that the multispin kernels have one code path
            CREATE_TRY(pool_alloc(s, (void **)&s->d_Jw, sizeof(uint32_t) * (size_t)(s->G * z * N)));
            CREATE_TRY(cudaMemsetAsync(s->d_Jw, 0, sizeof(uint32_t) * (size_t)(s->G * z * N), s->stream));
        }
        if (s->d_Jf) { pool_free(s, s->d_Jf); s->d_Jf = nullptr; }
        CREATE_TRY(pool_alloc(s, (void **)&s->d_words, sizeof(uint32_t) * (size_t)(s->G * m.S * N)));
    } else {
        if (m.coupling_class == COUP_UNIT) {
            CREATE_TRY(pool_alloc(s, (void **)&s->d_J8, (size_t)n_coup));
            couplings_to_int8_kernel<<<blocks_for(n_coup, 256), 256, 0, s->stream>>>(s->d_Jf, s->d_J8, n_coup);
            CREATE_TRY(cudaStreamSynchronize(s->stream));
        }
        if (m.coupling_class != COUP_F32 && s->d_Jf) { pool_free(s, s->d_Jf); s->d_Jf = nullptr; }
        CREATE_TRY(pool_alloc(s, (void **)&s->d_spins, (size_t)(m.D * m.S * N)));
        // per-row stride tables (pp_kernels_rows.cuh) whenever the colouring alternates along the rows
        const RowsPlan rp = getenv("PP_NO_ROWS") ? RowsPlan() : rows_plan(s->plan);
        if (rp.ok && N % 8 == 0 && z <= 16) {
            auto up = [&](const void *src, size_t bytes, const void **dst) -> cudaError_t {
                void *p = nullptr;
                cudaError_t e = pool_alloc(s, &p, bytes);
                if (e != cudaSuccess) return e;
                s->rows_bufs.push_back(p);
                *dst = p;
                return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
            };
            RowsView &v = s->rv;
            v.L = rp.L; v.kpr = rp.kpr; v.kpr_shift = rp.kpr_shift; v.m_half = rp.m_half; v.n_rows = rp.n_rows;
            CREATE_TRY(up(rp.row_a.data(), rp.row_a.size(), (const void **)&v.row_a));
            CREATE_TRY(up(rp.row_ord.data(), rp.row_ord.size() * 4, (const void **)&v.row_ord));
            CREATE_TRY(up(rp.nbr_row.data(), rp.nbr_row.size() * 4, (const void **)&v.nbr_row));
            CREATE_TRY(up(rp.dl.data(), rp.dl.size() * 4, (const void **)&v.dl));
            CREATE_TRY(up(rp.class_rows.data(), rp.class_rows.size() * 4, (const void **)&v.class_rows));
            CREATE_TRY(up(rp.class_start.data(), rp.class_start.size() * 4, (const void **)&v.class_start));
            CREATE_TRY(pool_alloc(s, (void **)&s->d_keys, sizeof(uint64_t) * (size_t)m.D));
            v.keys = s->d_keys;
            s->rows_class_start = rp.class_start;
            // split reductions: enough blocks to fill the GPU when there are few systems, each with >= 8 segments per thread
            const int64_t units = m.D * m.S, n_seg = rp.n_rows * rp.kpr;
            s->rows_nb = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, (4 * 148 + units - 1) / units), n_seg / (256 * 8)));
            const size_t n_acc = (size_t)std::max<int64_t>(units, m.D * (int64_t)(m.R / 2) * m.T);
            CREATE_TRY(pool_alloc(s, (void **)&s->d_rows_acc, sizeof(long long) * 2 * n_acc));
            CREATE_TRY(pool_alloc(s, (void **)&s->d_rows_arrive, sizeof(unsigned int) * n_acc));
            CREATE_TRY(cudaMemsetAsync(s->d_rows_acc, 0, sizeof(long long) * 2 * n_acc, s->stream));
            CREATE_TRY(cudaMemsetAsync(s->d_rows_arrive, 0, sizeof(unsigned int) * n_acc, s->stream));
            s->rows = true;
            // in-sweep energies on two-colour lattices (rows_sweep_kernel EACC).  fp32 couplings: per-thread bond sums are rounded
            // to integers in units of 1 / escale, a power of two chosen so that 4 sites * 2z' * max|J| * escale stays below 2^46
            // Worth it where a sweep is launch-bound (one launch less); large batches keep the streaming energy kernel (C4 at
            // D = 128: 155 attempts/ns with it, 151 with in-sweep energies).
            s->rows_esw = s->plan.n_colours == 2 && m.D * (int64_t)m.S * N <= (int64_t(1) << 24);
            if (const char *e = getenv("PP_ROWS_ESW")) s->rows_esw = s->plan.n_colours == 2 && atoi(e) != 0;
            if (s->rows_esw && m.coupling_class == COUP_F32 && desc->couplings) {
                float jmax = 0.0f;
                for (int64_t i = 0; i < n_coup; i++) jmax = std::max(jmax, std::fabs(desc->couplings[i]));
                int k = 30;
                if (jmax > 0.0f && std::isfinite(jmax)) k = std::min(30, (int)std::floor(std::log2(std::ldexp(1.0, 40) / ((double)z * jmax))));
                if (!std::isfinite(jmax)) s->rows_esw = false;
                s->rows_escale = std::ldexp(1.0f, k);
            }
            // resident kernel (rows_resident_kernel): integer classes, the realization's spins + the acceptance table in one
            // CTA's shared memory, and few enough segments per thread that one CTA is not slower than a grid
            {
                const size_t lut_words = ((size_t)m.T * (4 * z + 1) + 3) & ~size_t(3);
                const size_t need = lut_words * 4 + resident_scalar_bytes(m.S, m.T, m.P) + (size_t)m.S * N;
                s->resident_smem = need;
                s->resident = m.coupling_class != COUP_F32 && ((size_t)m.S * N) % 16 == 0 && need <= 200 * 1024 &&
                              (z == 2 || z == 3);
                if (const char *e = getenv("PP_RESIDENT")) s->resident = s->resident && atoi(e) != 0;
                // ferromagnets whose rows are whole 32-site words: the bit-packed form of the same kernel (pp_kernels_prows.cuh);
                // the handle then draws in the packed mapping whichever kernel runs a sweep
                bool dl1 = true;
                for (int k = 0; k < z; k++) dl1 = dl1 && std::abs(rp.dl[(size_t)k]) <= 1;
                s->resident_packed_smem = prows_resident_smem(m.S, m.T, m.P, z, N);
                s->resident_packed = s->resident && m.coupling_class == COUP_FERRO && rp.L % 32 == 0 && dl1 && (z == 2 || z == 3) &&
                                     s->resident_packed_smem <= 200 * 1024;
                if (const char *e = getenv("PP_RESIDENT_PACKED")) s->resident_packed = s->resident_packed && atoi(e) != 0;
                // ... spread over a thread-block cluster: two colours, one row class, the systems divide evenly, <= 2 work items per thread
                if (s->resident_packed && s->plan.n_colours == 2 && rp.m_half == 1 && !getenv("PP_NO_RESIDENT_CLUSTER")) {
                    const int64_t sysw = rp.n_rows * (rp.L / 32);
                    const int want_nc = getenv("PP_RESIDENT_CLUSTER") ? atoi(getenv("PP_RESIDENT_CLUSTER")) : 0;
                    for (int nc : {16, 8, 4, 2}) {
                        if (want_nc > 0 && nc != want_nc) continue;
                        if (nc == 16 && want_nc != 16) continue;  // non-portable cluster size: only on request (PP_RESIDENT_CLUSTER=16)
                        if (m.S % nc != 0 || m.S / nc > 64 || sysw > 65535) continue;
                        const int64_t items = (int64_t)(m.S / nc) * sysw;
                        const int nt = (int)std::min<int64_t>(512, std::max<int64_t>(64, (items + 31) / 32 * 32));
                        if (items > 2 * nt || nt < m.R) continue;
                        if (cluster_resident_layout(m.S, m.T, m.P, z, nc, (int)sysw).total_bytes > 200 * 1024) continue;
                        s->resident_cluster = nc;
                        s->resident_cluster_threads = nt;
                        break;
                    }
                }
                v.packed_draws = s->resident_packed ? 1 : 0;
            }
            // ferromagnets with one bit per spin, a system resident in one CTA's shared memory (pp_kernels_prows.cuh): rows that
            // split into whole 64-site word pairs, offsets that move by at most one site along the rows, 2-4 forward directions
            {
                bool dl_ok = true;
                for (int k = 0; k < z; k++) dl_ok = dl_ok && std::abs(rp.dl[(size_t)k]) <= 1;
                const int64_t W = rp.L / 64, sys_words = rp.n_rows * 2 * W;
                s->prows = m.coupling_class == COUP_FERRO && rp.L % 64 == 0 && dl_ok && z >= 2 && z <= PROWS_MAX_Z &&
                           sys_words * 4 <= 160 * 1024 && !s->resident;
                if (const char *e = getenv("PP_ROWS_PACKED")) s->prows = s->prows && atoi(e) != 0;
                if (s->prows) {
                    s->prows_cluster = m.R >= 1 && m.R <= 8 && sys_words * 8 <= 160 * 1024 && sys_ranks == 1 && !getenv("PP_PROWS_NO_CLUSTER");
                    s->pv.W = (int)W;
                    s->pv.sys_words = sys_words;
                    CREATE_TRY(pool_alloc(s, (void **)&s->pv.words, sizeof(uint32_t) * (size_t)(m.D * m.S * sys_words)));
                }
            }
            // fp32 couplings: the same site of 32 systems of a realization in one word (pp_kernels_swords.cuh)
            {
                bool dl_ok = true;
                for (int k = 0; k < z; k++) dl_ok = dl_ok && std::abs(rp.dl[(size_t)k]) <= 1;
                float jmax = 0.0f;
                if (m.coupling_class == COUP_F32 && desc->couplings)
                    for (int64_t i = 0; i < n_coup; i++) jmax = std::max(jmax, std::fabs(desc->couplings[i]));
                s->swords = m.coupling_class == COUP_F32 && desc->couplings && s->plan.n_colours == 2 && rp.m_half == 1 && (rp.L % 32 == 0 || (rp.L < 32 && N % 32 == 0)) &&
                            dl_ok && (z == 2 || z == 3) && m.S >= 16 && sys_ranks == 1 && std::isfinite(jmax) && jmax > 0.0f &&
                            N / 8 <= 160 * 1024 && N < (int64_t(1) << 28);
                if (const char *e = getenv("PP_SYS_WORDS")) s->swords = s->swords && atoi(e) != 0;
                if (s->swords) {
                    SWordsView &sv = s->swv;
                    sv.SW = (m.S + 31) / 32;
                    // fixed-point unit of the in-sweep bond sums: |s h| escale <= 2z' max|J| escale stays below 2^20 (the rounding addend
                    // 1.5 * 2^23 holds 2^22; a thread adds SW_SPT terms, a warp 32 threads, a block four warps into 32-bit integers)
                    sv.escale = std::ldexp(1.0f, std::min(30, (int)std::floor(std::log2(std::ldexp(1.0, 20) / (2.0 * z * (double)jmax)))));
                    const size_t n_acc = (size_t)m.D * (size_t)sv.SW;
                    auto take = [&](void **p, size_t bytes, bool zero) -> cudaError_t {
                        cudaError_t e = pool_alloc(s, p, bytes);
                        if (e != cudaSuccess) return e;
                        s->rows_bufs.push_back(*p);
                        return zero ? cudaMemsetAsync(*p, 0, bytes, s->stream) : cudaSuccess;
                    };
                    CREATE_TRY(take((void **)&sv.words, sizeof(uint32_t) * n_acc * (size_t)N, true));
                    CREATE_TRY(take((void **)&sv.tbits, sizeof(uint32_t) * (size_t)m.D * (size_t)m.S * (size_t)(N / 32), false));
                    CREATE_TRY(take((void **)&sv.acc_e, sizeof(long long) * n_acc * 32, true));
                    CREATE_TRY(take((void **)&sv.acc_m, sizeof(long long) * n_acc * 32, true));
                    CREATE_TRY(take((void **)&sv.arrive_e, sizeof(unsigned int) * n_acc, true));
                    CREATE_TRY(take((void **)&sv.arrive_m, sizeof(unsigned int) * n_acc, true));
                    CREATE_TRY(take((void **)&sv.lane_t, sizeof(float4) * n_acc * 32, true));
                }
            }
        }
    }
    if (sys_ranks > 1) {
        if (s->layout != PP_LAYOUT_INT8 || !s->rows) {
            free_sim(s);
            return fail(PP_ERR_UNSUPPORTED, "system_ranks > 1 needs the row-table int8 kernels (a linear colouring that alternates along the "
                                            "rows, last extent a multiple of 8)");
        }
        s->resident = false;  // one CTA per realization would hold every system (the draw mapping of the handle stays what it was)
        s->resident_packed = false;
        s->sys_ranks = sys_ranks;
        s->sys_rank = desc->system_rank;
        const int per = m.S / sys_ranks;
        m.sys_lo = s->sys_rank * per;
        m.sys_hi = m.sys_lo + per;
        pp_status cst = cached_comm(s->device, sys_ranks, s->sys_rank, desc->nccl_unique_id, &s->sys_comm);
        if (cst != PP_OK) {
            std::string keep = g_last_error;
            free_sim(s);
            return fail(cst, keep);
        }
    }
    m.J8 = s->d_J8;
    m.Jf = s->d_Jf;
    m.Jw = s->d_Jw;
    m.spins = s->d_spins;
    m.words = s->d_words;

    const int64_t DS = m.D * m.S;
    CREATE_TRY(pool_alloc(s, (void **)&s->d_sid, sizeof(int32_t) * (size_t)DS));
    CREATE_TRY(pool_alloc(s, (void **)&s->d_energies, sizeof(float) * (size_t)DS));
    CREATE_TRY(pool_alloc(s, (void **)&s->d_mags, sizeof(long long) * (size_t)DS));
    CREATE_TRY(cudaMemset(s->d_mags, 0, sizeof(long long) * (size_t)DS));
    CREATE_TRY(pool_alloc(s, (void **)&s->d_temps, sizeof(float) * (size_t)m.T));
    CREATE_TRY(cudaMemcpy(s->d_temps, s->temps.data(), sizeof(float) * (size_t)m.T, cudaMemcpyHostToDevice));
    m.system_ids = s->d_sid;
    m.energies = s->d_energies;
    m.mags = s->d_mags;
    m.temps = s->d_temps;

    if (m.coupling_class != COUP_F32) {
        const int width = 4 * z + 1;
        std::vector<uint32_t> lut((size_t)m.T * width);
        for (int mode = 0; mode < 2; mode++) {
            pp_metropolis_lookup(s->temps.data(), m.T, z, mode, lut.data());
            uint32_t *&dst = mode == 0 ? s->d_lut_metro : s->d_lut_gibbs;
            CREATE_TRY(pool_alloc(s, (void **)&dst, sizeof(uint32_t) * lut.size()));
            CREATE_TRY(cudaMemcpy(dst, lut.data(), sizeof(uint32_t) * lut.size(), cudaMemcpyHostToDevice));
        }
    }

    // packed-row kernel: thresholds compared per site -- z' when the counts for unsat >= z' are 2^24 (energy change <= 0 always
    // accepted) and the others are below 2^24, else all 2z' + 1
    if (s->prows || s->resident_packed) {
        std::vector<uint32_t> lut((size_t)m.T * (4 * z + 1));
        for (int mode = 0; mode < 2; mode++) {
            pp_metropolis_lookup(s->temps.data(), m.T, z, mode, lut.data());
            bool fast = true, bs = z == 3;
            for (int t = 0; t < m.T; t++)
                for (int u = 0; u <= 2 * z; u++) {
                    const uint32_t c = lut[(size_t)t * (4 * z + 1) + 2 * u];
                    if (u >= z ? c != F24 : c >= F24) fast = false;
                    if (c >= F24 || (u > 0 && c < lut[(size_t)t * (4 * z + 1) + 2 * (u - 1)])) bs = false;
                }
            s->prows_nm[mode] = fast ? z : 2 * z + 1;
            s->prows_bs[mode] = bs && !fast && !getenv("PP_PROWS_NO_BS");
        }
    }

    // bit-packed slab kernel: how many thresholds it compares per site, and whether the counts grow with unsat (they do for
    // every finite positive temperature; checked because the kernel's carry chain relies on it)
    if (s->layout == PP_LAYOUT_SLAB && s->slab->packed) {
        SlabState *sl = s->slab;
        std::vector<uint32_t> lut((size_t)m.T * 13);
        for (int mode = 0; mode < 2; mode++) {
            pp_metropolis_lookup(s->temps.data(), m.T, 3, mode, lut.data());
            bool mono = true, fast = true;
            for (int t = 0; t < m.T; t++)
                for (int u = 0; u <= 6; u++) {
                    const uint32_t c = lut[(size_t)t * 13 + 2 * u];
                    if (u > 0 && c < lut[(size_t)t * 13 + 2 * (u - 1)]) mono = false;
                    if (u >= 3 ? c != F24 : c >= F24) fast = false;
                }
            (mode == 0 ? sl->mono_metro : sl->mono_gibbs) = mono;
            (mode == 0 ? sl->nm_metro : sl->nm_gibbs) = fast ? 3 : 7;
        }
    }

    // specialised 3-D multispin kernel (pp_kernels_msc3d.cuh)
    if (s->layout == PP_LAYOUT_MSC && z == 3 && (m.R == 1 || m.R == 2 || m.R == 4)) {
        s->m3 = msc3d_plan(s->plan);
        if (s->m3.ok) {
            // NH = 1: one temperature slot per CTA, preferred when two such CTAs fit one SM (they run out of phase, so one
            // stages data while the other computes); NH = 2: two slots per CTA sharing the coupling words
            auto smem_words = [&](int nh) {
                const size_t jw = 3 * (size_t)N;
                return nh == 2 ? jw + 4 * (size_t)s->m3.n_items + 2 * (size_t)m.R * N + 2 * MSC3D_NBAR + 2 * 64 + 2 * 512 : jw + (size_t)m.R * N + 2 * MSC3D_NBAR + 64;  // 64 words per half: parked system ids
            };
            const size_t sm_total = 227 * 1024, cta_reserved = 1024;
            int nh = 1;
            if (2 * (smem_words(1) * 4 + cta_reserved) > sm_total && smem_words(2) * 4 + cta_reserved <= sm_total) nh = 2;
            if ((size_t)m.R * N < 1152) nh = 2;  // NH = 1 parks its 2 KB reduction scratch + 2.3 KB of tail values in the spin buffer
            if (const char *e = getenv("PP_MSC3D_NH")) nh = atoi(e) == 2 ? 2 : 1;
            // per-thread counter capacity of the epilogue (MSC3D_KE / MSC3D_KM planes)
            const int64_t sites_em = N / (32 * (4 / m.R)), sites_pair = m.P > 0 ? N / (32 * (4 / m.P)) : 0;
            const bool cap_ok = 3 * sites_em < (1 << MSC3D_KE) - 8 && sites_em < (1 << MSC3D_KM) - 8 &&
                                3 * sites_pair < (1 << MSC3D_KE) - 8;
            if (cap_ok && smem_words(nh) * 4 + cta_reserved <= sm_total && (nh == 2 || (size_t)m.R * N >= 1152)) {
                CREATE_TRY(pool_alloc(s, (void **)&s->d_items, sizeof(uint16_t) * s->m3.items.size()));
                CREATE_TRY(cudaMemcpy(s->d_items, s->m3.items.data(), sizeof(uint16_t) * s->m3.items.size(), cudaMemcpyHostToDevice));
                s->gv.items = s->d_items;
                s->gv.n_items = s->m3.n_items;
                s->gv.N = (uint32_t)N;
                s->gv.N2 = (uint32_t)(N / 2);
                s->gv.one[0] = s->gv.one[1] = s->gv.one[2] = 1u;
                s->msc3d = true;
                CREATE_TRY(pool_alloc(s, (void **)&s->d_words_alt, sizeof(uint32_t) * (size_t)(s->G * m.S * N)));
                s->msc3d_nh = nh;
                s->msc3d_smem = smem_words(nh) * 4;
                // in-sweep energy: one slot per CTA, replicas in pairs, at most two quads per thread (MSC3D_KS planes) and
                // at most four pair items per thread (MSC3D_KQ / MSC3D_KL planes); the coupling-word area must hold the
                // parked counters (R * KS * 256 words, then 25 planes * 256 words) below its last 512 words
                {
                    const int64_t park = std::max<int64_t>((int64_t)m.R * MSC3D_KS * MSC3D_NTH, (3 * MSC3D_KQ + MSC3D_KL) * MSC3D_NTH);
                    s->msc3d_esw = nh == 1 && (m.R == 2 || m.R == 4) && s->m3.n_items <= 2 * MSC3D_NTH &&
                                   3 * (int64_t)N - 512 >= park;
                    if (const char *e = getenv("PP_MSC3D_ESW")) s->msc3d_esw = s->msc3d_esw && atoi(e) != 0;
                }
                // Metropolis fast path: counts for unsat >= 3 (energy_change >= 0) are 2^24 (sweep.rs:141-145) and the
                // others are below 2^24, so (draw < count) == (raw32 < count << 8)
                std::vector<uint32_t> lut((size_t)m.T * 13);
                pp_metropolis_lookup(s->temps.data(), m.T, 3, PP_SWEEP_METROPOLIS, lut.data());
                s->msc3d_metro = true;
                for (int t = 0; t < m.T; t++)
                    for (int u = 0; u <= 6; u++) {
                        const uint32_t c = lut[(size_t)t * 13 + 2 * u];
                        if (u >= 3 ? c != F24 : c >= F24) s->msc3d_metro = false;
                    }
            }
        }
    }

    // PT state (realization.rs:21-67)
    const int n_edges = m.T - 1;
    CREATE_TRY(pool_alloc(s, (void **)&s->pt.edge_attempts, sizeof(uint64_t) * (size_t)(m.D * std::max(n_edges, 1))));
    CREATE_TRY(pool_alloc(s, (void **)&s->pt.edge_acceptances, sizeof(uint64_t) * (size_t)(m.D * std::max(n_edges, 1))));
    CREATE_TRY(pool_alloc(s, (void **)&s->pt.round_trips, sizeof(uint64_t) * (size_t)DS));
    CREATE_TRY(pool_alloc(s, (void **)&s->pt.trip_state, (size_t)DS));
    if (s->layout == PP_LAYOUT_MSC)
        CREATE_TRY(pool_alloc(s, (void **)&s->pt.swap_mask, sizeof(uint32_t) * (size_t)(s->G * m.R * std::max(n_edges, 1))));
    s->pt.cold_slot = s->pt.hot_slot = 0;  // realization.rs:92-107
    for (int slot = 1; slot < m.T; slot++) {
        if (s->temps[slot] < s->temps[s->pt.cold_slot]) s->pt.cold_slot = slot;
        if (s->temps[slot] > s->temps[s->pt.hot_slot]) s->pt.hot_slot = slot;
    }

    // stats
    CREATE_TRY(pool_alloc(s, (void **)&s->st.sums, sizeof(double) * (size_t)(m.D * 11 * m.T)));
    if (m.P > 0) {
        CREATE_TRY(pool_alloc(s, (void **)&s->d_dot_spin, sizeof(long long) * (size_t)(m.D * m.P * m.T)));
        CREATE_TRY(pool_alloc(s, (void **)&s->d_dot_link, sizeof(long long) * (size_t)(m.D * m.P * m.T)));
        s->st.dot_spin = s->d_dot_spin;
        s->st.dot_link = s->d_dot_link;
    }
#undef CREATE_TRY
    pp_status st = do_reset(s, desc->seed);
    if (st != PP_OK) {
        std::string keep = g_last_error;
        free_sim(s);
        return fail(st, keep);
    }
    *out = s;
    return PP_OK;
}

extern "C" pp_status pp_reset(pp_sim *sim, int32_t has_seed, uint64_t seed) {
    if (!sim) return fail(PP_ERR_INVALID, "sim is NULL");
    return do_reset(sim, has_seed ? seed : sim->ctor_seed);  // lib.rs:626
}

// config.rs:180-247 (messages kept: the reference's tests match on them)
static pp_status validate_cfg(const pp_sample_cfg *c) {
    if (c->n_sweeps < 1) return fail(PP_ERR_INVALID, "n_sweeps must be >= 1");
    if (c->warmup_sweeps > c->n_sweeps || c->warmup_sweeps < 0) return fail(PP_ERR_INVALID, "warmup_sweeps must be <= n_sweeps");
    if (c->pt_interval < 0) return fail(PP_ERR_INVALID, "pt_interval must be >= 1");
    if (c->sweep_mode != PP_SWEEP_METROPOLIS && c->sweep_mode != PP_SWEEP_GIBBS)
        return fail(PP_ERR_INVALID, "unknown sweep_mode, expected 'metropolis' or 'gibbs'");
    if (c->pt_schedule != PP_PT_SINGLE_RANDOM_EDGE && c->pt_schedule != PP_PT_FULL_LADDER)
        return fail(PP_ERR_INVALID, "unknown pt_schedule, expected 'single_random_edge' or 'full_ladder'");
    if (c->cluster_update_interval < 0) return fail(PP_ERR_INVALID, "cluster_update_interval must be >= 1");
    if (c->cluster_mode != PP_CLUSTER_SW && c->cluster_mode != PP_CLUSTER_WOLFF)
        return fail(PP_ERR_INVALID, "unknown cluster_mode, expected 'wolff' or 'sw'");
    if (c->overlap_cluster_update_interval < 0) return fail(PP_ERR_INVALID, "overlap_cluster_update_interval must be >= 1");
    if (c->overlap_cluster_mode != PP_CLUSTER_SW && c->overlap_cluster_mode != PP_CLUSTER_WOLFF)
        return fail(PP_ERR_INVALID, "unknown cluster_mode, expected 'wolff' or 'sw'");
    if (c->autocorrelation_max_lag < 0) return fail(PP_ERR_INVALID, "autocorrelation_max_lag must be >= 1");
    if (c->snapshot_interval != 0) return fail(PP_ERR_UNSUPPORTED, "snapshot_interval is not implemented on the GPU sweep path");
    return PP_OK;
}

static pp_status ensure_tables(pp_sim *s, bool need_log, bool need_glog) {
    if (need_log && !s->mv.logtab) {
        pp_status st = get_log_table(s->device, false, &s->mv.logtab);
        if (st != PP_OK) return st;
    }
    if (need_glog && !s->mv.glogtab) {
        pp_status st = get_log_table(s->device, true, &s->mv.glogtab);
        if (st != PP_OK) return st;
    }
    return PP_OK;
}

static pp_status ensure_streams(pp_sim *s, int n) {
    while ((int)s->xstreams.size() < n) {
        cudaStream_t x;
        cudaEvent_t e;
        CUDA_TRY(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        s->xstreams.push_back(x);
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->xevents.push_back(e);
    }
    return PP_OK;
}

static pp_status ensure_hist(pp_sim *s) {
    if (s->hist_allocated || s->mv.P == 0) return PP_OK;
    const size_t n = (size_t)s->mv.D * s->mv.T * (s->mv.N + 1);
    CUDA_TRY(pool_alloc(s, (void **)&s->st.hist, sizeof(uint32_t) * n));
    CUDA_TRY(pool_alloc(s, (void **)&s->st.ql_at_q, sizeof(double) * n));
    CUDA_TRY(pool_alloc(s, (void **)&s->st.ql2_at_q, sizeof(double) * n));
    s->hist_allocated = true;
    return PP_OK;
}

// Small realizations (S * N bytes + the acceptance table + the scalar state fit one CTA's shared memory, integer coupling
// classes): every sweep of the call inside rows_resident_kernel, up to 256 sweeps per launch.
static pp_status run_rows_resident(pp_sim *s, Ctx &c, const pp_sample_cfg *cfg, const volatile int32_t *interrupt,
                                   void (*on_sweep)(void *, uint64_t), void *user) {
    int64_t sweep_id = 0;
    RowsView v = s->rv;
    v.keys = s->d_keys;
    ModelView mk = c.m;
    mk.lut = cfg->sweep_mode == PP_SWEEP_GIBBS ? s->d_lut_gibbs : s->d_lut_metro;
    const bool gibbs = cfg->sweep_mode == PP_SWEEP_GIBBS;
    while (sweep_id < cfg->n_sweeps) {
        if (interrupt && *interrupt) return fail(PP_ERR_INTERRUPTED, "interrupted");
        const int64_t mb_end = std::min<int64_t>(cfg->n_sweeps, sweep_id + 256);
        ResidentArgs a;
        a.sweep_id0 = sweep_id;
        a.n_sweeps = (int)(mb_end - sweep_id);
        a.warmup_sweeps = cfg->warmup_sweeps;
        a.pt_interval = cfg->pt_interval > 0 ? cfg->pt_interval : 0;
        a.pt_schedule = cfg->pt_schedule == PP_PT_FULL_LADDER ? 1 : 0;
        a.sweep_counter0 = s->sweep_counter;
        a.pt_event0 = s->pt_event_counter;
        a.parity0 = s->next_parity;
        a.spins_in_smem = 1;
        a.dot_spin = c.dot_spin;
        a.dot_link = c.dot_link;
#define PP_RES3(C_, Z_, G_)                                                                                                       \
    do {                                                                                                                          \
    CUDA_TRY(cudaFuncSetAttribute(rows_resident_kernel<C_, Z_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->resident_smem)); \
    rows_resident_kernel<C_, Z_, G_><<<(unsigned)mk.D, RESIDENT_THREADS, s->resident_smem, c.stream>>>(mk, v, c.st, c.pt, a);   \
    } while (0)
#define PP_RES2(C_, Z_) do { if (gibbs) PP_RES3(C_, Z_, true); else PP_RES3(C_, Z_, false); } while (0)
#define PP_PRES(Z_, NM_)                                                                                                           \
    do {                                                                                                                          \
    CUDA_TRY(cudaFuncSetAttribute(prows_resident_kernel<Z_, NM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->resident_packed_smem)); \
    prows_resident_kernel<Z_, NM_><<<(unsigned)mk.D, PRES_THREADS, s->resident_packed_smem, c.stream>>>(mk, v, c.st, c.pt, a);      \
    } while (0)
        if (s->resident_packed && s->resident_cluster > 0) {
            const int nm = s->prows_nm[gibbs ? 1 : 0], nc = s->resident_cluster;
            const size_t smem = cluster_resident_layout(mk.S, mk.T, mk.P, mk.z, nc, (int)(v.n_rows * (v.L / 32))).total_bytes;
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3((unsigned)(mk.D * nc));
            lc.blockDim = dim3((unsigned)s->resident_cluster_threads);
            lc.dynamicSmemBytes = smem;
            lc.stream = c.stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)nc;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            lc.attrs = attr;
            lc.numAttrs = 1;
#define PP_CRES(Z_, NM_)                                                                                                              \
    do {                                                                                                                              \
    CUDA_TRY(cudaFuncSetAttribute(prows_cluster_resident_kernel<Z_, NM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    if (nc > 8) CUDA_TRY(cudaFuncSetAttribute(prows_cluster_resident_kernel<Z_, NM_>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
    CUDA_TRY(cudaLaunchKernelEx(&lc, prows_cluster_resident_kernel<Z_, NM_>, mk, v, c.st, c.pt, a, nc));                               \
    } while (0)
            if (mk.z == 2) { if (nm == 2) PP_CRES(2, 2); else PP_CRES(2, 5); }
            else { if (nm == 3) PP_CRES(3, 3); else PP_CRES(3, 7); }
#undef PP_CRES
        } else if (s->resident_packed) {
            const int nm = s->prows_nm[gibbs ? 1 : 0];
            if (mk.z == 2) { if (nm == 2) PP_PRES(2, 2); else PP_PRES(2, 5); }
            else { if (nm == 3) PP_PRES(3, 3); else PP_PRES(3, 7); }
        } else if (mk.coupling_class == COUP_FERRO) { if (mk.z == 2) PP_RES2(COUP_FERRO, 2); else PP_RES2(COUP_FERRO, 3); }
        else { if (mk.z == 2) PP_RES2(COUP_UNIT, 2); else PP_RES2(COUP_UNIT, 3); }
#undef PP_PRES
#undef PP_RES2
#undef PP_RES3
        s->launches++;
        CUDA_TRY(cudaGetLastError());
        for (int64_t sid = sweep_id; sid < mb_end; sid++) {
            s->sweep_counter++;
            if (cfg->pt_interval > 0 && sid % cfg->pt_interval == 0) {
                s->pt_event_counter++;
                if (mk.T >= 2 && cfg->pt_schedule == PP_PT_FULL_LADDER) s->next_parity = 1 - s->next_parity;
            }
            if (on_sweep) on_sweep(user, (uint64_t)sid);
        }
        sweep_id = mb_end;
    }
    return PP_OK;
}

// Device accumulators -> the caller's result arrays (statistics/stats.rs:29-35, results.rs:165-180, 250-259, overlap.rs:106-152):
// per-realization means, their ordered sum over realizations, histograms, taus, equilibration checkpoints, PT counters.
static pp_status collect_results(pp_sim *s, pp_results *out, int64_t n_rec, bool want_eq, bool want_ac, const std::vector<int64_t> &ckpts,
                                 const std::vector<double> &eq_snap, const std::vector<double> &taus) {
    const ModelView &m = s->mv;
    const int T = m.T;
    std::vector<double> sums((size_t)m.D * 11 * T);
    CUDA_TRY(cudaMemcpy(sums.data(), s->st.sums, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost));
    const double c_stat = n_rec > 0 ? (double)(n_rec * m.R) : 1.0;
    const double c_ov = (n_rec > 0 && m.P > 0) ? (double)(n_rec * m.P) : 1.0;
    // one pass in storage order; every (k, t) accumulator still adds its realizations in order (results.rs:165-180)
    std::vector<double> mean_acc((size_t)11 * T, 0.0);
    for (int64_t d = 0; d < m.D; d++)
        for (int k = 0; k < 11; k++) {
            const double c = k < 5 ? c_stat : c_ov;
            double *row = &sums[((size_t)d * 11 + k) * T], *acc = &mean_acc[(size_t)k * T];
            for (int t = 0; t < T; t++) {
                row[t] /= c;
                acc[t] += row[t];
            }
        }
    if (out->per_sample_means) memcpy(out->per_sample_means, sums.data(), sizeof(double) * sums.size());
    double *dst[11] = {out->mags, out->mags2, out->mags4, out->energies, out->energies2, out->overlap, out->overlap2,
                       out->overlap4, out->link_overlap, out->link_overlap2, out->link_overlap4};
    for (int k = 0; k < 11; k++) {
        if (!dst[k] || (k >= 5 && m.P == 0)) continue;
        for (int t = 0; t < T; t++) dst[k][t] = mean_acc[(size_t)k * T + t] / (double)m.D;
    }
    if (want_eq) {  // results.rs:231-247, 275-282
        double *edst[2] = {out->equil_energy_avg, out->equil_link_overlap_avg};
        const size_t nck = ckpts.size();
        for (int k = 0; k < 2; k++) {
            if (!edst[k]) continue;
            for (size_t cidx = 0; cidx < nck; cidx++)
                for (int t = 0; t < T; t++) {
                    double acc = 0.0;
                    for (int64_t d = 0; d < m.D; d++) acc += eq_snap[(((size_t)d * nck + cidx) * 2 + k) * T + t];
                    edst[k][cidx * T + t] = acc / (double)m.D;
                }
        }
        if (out->per_sample_equil) memcpy(out->per_sample_equil, eq_snap.data(), sizeof(double) * eq_snap.size());
    }
    if (want_ac) {  // results.rs:217-231, 269-274: sum over realizations in order, divide by their number
        double *tdst[2] = {out->mags2_tau, m.P > 0 ? out->overlap2_tau : nullptr};
        for (int k = 0; k < 2; k++) {
            if (!tdst[k]) continue;
            for (int t = 0; t < T; t++) {
                double acc = 0.0;
                for (int64_t d = 0; d < m.D; d++) acc += taus[(size_t)k * m.D * T + (size_t)d * T + t];
                tdst[k][t] = acc / (double)m.D;
            }
        }
        if (out->per_sample_taus)
            for (int64_t d = 0; d < m.D; d++)
                for (int k = 0; k < 2; k++)
                    memcpy(out->per_sample_taus + ((size_t)d * 2 + k) * T, &taus[(size_t)k * m.D * T + (size_t)d * T], sizeof(double) * (size_t)T);
    }
    if (m.P > 0) {
        const int64_t per = (int64_t)T * (m.N + 1);
        if (out->overlap_histogram || out->ql_at_q_sum || out->ql2_at_q_sum) {
            if (!s->hist_allocated) {
                if (out->overlap_histogram) memset(out->overlap_histogram, 0, sizeof(uint64_t) * (size_t)per);
                if (out->ql_at_q_sum) memset(out->ql_at_q_sum, 0, sizeof(double) * (size_t)per);
                if (out->ql2_at_q_sum) memset(out->ql2_at_q_sum, 0, sizeof(double) * (size_t)per);
            } else {
                unsigned long long *d_h = nullptr;
                double *d_a = nullptr, *d_b = nullptr;
                CUDA_TRY(pool_alloc(s, (void **)&d_h, sizeof(uint64_t) * (size_t)per));
                CUDA_TRY(pool_alloc(s, (void **)&d_a, sizeof(double) * (size_t)per));
                CUDA_TRY(pool_alloc(s, (void **)&d_b, sizeof(double) * (size_t)per));
                reduce_hist_kernel<<<blocks_for(per, 256), 256, 0, s->stream>>>(m, s->st, d_h, d_a, d_b);
                CUDA_TRY(cudaStreamSynchronize(s->stream));
                if (out->overlap_histogram) CUDA_TRY(cudaMemcpy(out->overlap_histogram, d_h, sizeof(uint64_t) * (size_t)per, cudaMemcpyDeviceToHost));
                if (out->ql_at_q_sum) CUDA_TRY(cudaMemcpy(out->ql_at_q_sum, d_a, sizeof(double) * (size_t)per, cudaMemcpyDeviceToHost));
                if (out->ql2_at_q_sum) CUDA_TRY(cudaMemcpy(out->ql2_at_q_sum, d_b, sizeof(double) * (size_t)per, cudaMemcpyDeviceToHost));
                pool_free(s, d_h); pool_free(s, d_a); pool_free(s, d_b);
            }
        }
        const size_t all = (size_t)m.D * per;
        if (out->per_sample_overlap_histogram) {
            if (!s->hist_allocated) memset(out->per_sample_overlap_histogram, 0, sizeof(uint64_t) * all);
            else {
                const int64_t chunk = std::min<int64_t>((int64_t)all, int64_t(1) << 26);
                unsigned long long *d_w = nullptr;
                CUDA_TRY(pool_alloc(s, (void **)&d_w, sizeof(uint64_t) * (size_t)chunk));
                for (int64_t off = 0; off < (int64_t)all; off += chunk) {
                    const int64_t n = std::min<int64_t>(chunk, (int64_t)all - off);
                    widen_u32_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(s->st.hist + off, d_w, n);
                    CUDA_TRY(cudaStreamSynchronize(s->stream));
                    CUDA_TRY(cudaMemcpy(out->per_sample_overlap_histogram + off, d_w, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost));
                }
                pool_free(s, d_w);
            }
        }
        if (out->per_sample_ql_at_q_sum) {
            if (!s->hist_allocated) memset(out->per_sample_ql_at_q_sum, 0, sizeof(double) * all);
            else CUDA_TRY(cudaMemcpy(out->per_sample_ql_at_q_sum, s->st.ql_at_q, sizeof(double) * all, cudaMemcpyDeviceToHost));
        }
        if (out->per_sample_ql2_at_q_sum) {
            if (!s->hist_allocated) memset(out->per_sample_ql2_at_q_sum, 0, sizeof(double) * all);
            else CUDA_TRY(cudaMemcpy(out->per_sample_ql2_at_q_sum, s->st.ql2_at_q, sizeof(double) * all, cudaMemcpyDeviceToHost));
        }
    }
    if (T > 1) {
        if (out->pt_edge_attempts)
            CUDA_TRY(cudaMemcpy(out->pt_edge_attempts, s->pt.edge_attempts, sizeof(uint64_t) * (size_t)(m.D * (T - 1)), cudaMemcpyDeviceToHost));
        if (out->pt_edge_acceptances)
            CUDA_TRY(cudaMemcpy(out->pt_edge_acceptances, s->pt.edge_acceptances, sizeof(uint64_t) * (size_t)(m.D * (T - 1)), cudaMemcpyDeviceToHost));
    }
    if (out->pt_round_trips)
        CUDA_TRY(cudaMemcpy(out->pt_round_trips, s->pt.round_trips, sizeof(uint64_t) * (size_t)(m.D * m.S), cudaMemcpyDeviceToHost));
    return PP_OK;
}

extern "C" pp_status pp_sample(pp_sim *s, const pp_sample_cfg *cfg, pp_results *out, const volatile int32_t *interrupt,
                               void (*on_sweep)(void *, uint64_t), void *user) {
    if (!s || !cfg) return fail(PP_ERR_INVALID, "sim/cfg is NULL");
    pp_status st = validate_cfg(cfg);  // before any mutation (tests/test_sampling_interfaces.py:145-156)
    if (st != PP_OK) return st;
    CUDA_TRY(cudaSetDevice(s->device));
    ModelView &m = s->mv;
    const bool f32 = m.coupling_class == COUP_F32;
    st = ensure_tables(s, cfg->pt_interval > 0 || (f32 && cfg->exact_log && cfg->sweep_mode == PP_SWEEP_METROPOLIS),
                       f32 && cfg->exact_log && cfg->sweep_mode == PP_SWEEP_GIBBS);
    if (st != PP_OK) return st;
    const int64_t n_rec = cfg->n_sweeps - cfg->warmup_sweeps;
    if (n_rec > 0) {
        st = ensure_hist(s);
        if (st != PP_OK) return st;
    }
    // fresh accumulators per call (simulation/mod.rs:331-335: Statistics::new per run)
    CUDA_TRY(cudaMemsetAsync(s->st.sums, 0, sizeof(double) * (size_t)(m.D * 11 * m.T), s->stream));
    if (s->hist_allocated) {
        const size_t n = (size_t)m.D * m.T * (m.N + 1);
        CUDA_TRY(cudaMemsetAsync(s->st.hist, 0, sizeof(uint32_t) * n, s->stream));
        CUDA_TRY(cudaMemsetAsync(s->st.ql_at_q, 0, sizeof(double) * n, s->stream));
        CUDA_TRY(cudaMemsetAsync(s->st.ql2_at_q, 0, sizeof(double) * n, s->stream));
    }
    // autocorrelation accumulators (mod.rs:341-371): fresh per call, one per (realization, temperature)
    AutocorrView ac_m{0, nullptr, nullptr, nullptr, nullptr}, ac_q{0, nullptr, nullptr, nullptr, nullptr};
    double *d_tau = nullptr;
    const bool want_ac = cfg->autocorrelation_max_lag > 0;
    auto free_ac = [&]() {
        pool_free(s, ac_m.ring); pool_free(s, ac_m.sum_prod); pool_free(s, ac_m.sum_o);
        pool_free(s, ac_q.ring); pool_free(s, ac_q.sum_prod); pool_free(s, ac_q.sum_o);
        pool_free(s, d_tau);
        ac_m = ac_q = AutocorrView{0, nullptr, nullptr, nullptr, nullptr};
        d_tau = nullptr;
    };
    // Every exit of this function runs `cleanup` (scope guard): per-call buffers go back to the pool, and once the sweep loop
    // has started a pending multispin exchange is applied and the handle's word pointer is put on the current ping-pong buffer,
    // so that an error or an interrupt never leaves the handle between two states.
    struct ScopeExit {
        std::function<void()> f;
        ~ScopeExit() { if (f) f(); }
    } cleanup;
    if (want_ac) {
        if (s->layout == PP_LAYOUT_SLAB) return fail(PP_ERR_UNSUPPORTED, "autocorrelation_max_lag is not implemented for the slab layout");
        const int K = (int)std::max<int64_t>(1, std::min<int64_t>(cfg->autocorrelation_max_lag, n_rec / 4));
        const size_t ndt = (size_t)m.D * m.T, len = (size_t)K + 1;
        for (AutocorrView *a : {&ac_m, &ac_q}) {
            if (a == &ac_q && m.P == 0) break;
            a->K = K;
            CUDA_TRY(pool_alloc(s, (void **)&a->ring, sizeof(float) * ndt * len));
            CUDA_TRY(pool_alloc(s, (void **)&a->sum_prod, sizeof(double) * ndt * len));
            CUDA_TRY(pool_alloc(s, (void **)&a->sum_o, sizeof(double) * ndt * 2));
            a->sum_o2 = a->sum_o + ndt;
            CUDA_TRY(cudaMemsetAsync(a->ring, 0, sizeof(float) * ndt * len, s->stream));
            CUDA_TRY(cudaMemsetAsync(a->sum_prod, 0, sizeof(double) * ndt * len, s->stream));
            CUDA_TRY(cudaMemsetAsync(a->sum_o, 0, sizeof(double) * ndt * 2, s->stream));
        }
        CUDA_TRY(pool_alloc(s, (void **)&d_tau, sizeof(double) * ndt * 2));
    }
    // equilibration diagnostic (mod.rs:373-383): running sums per (realization, temperature) + checkpoint snapshots
    const bool want_eq = cfg->equilibration_diagnostic != 0;
    std::vector<int64_t> ckpts;
    double *d_eq_sum = nullptr, *d_eq_snap = nullptr;
    auto free_eq = [&]() { pool_free(s, d_eq_sum); pool_free(s, d_eq_snap); d_eq_sum = d_eq_snap = nullptr; };
    uint32_t *d_fk_count = nullptr, *d_fk_lab = nullptr;
    uint8_t *d_fk_bm = nullptr;
    auto free_fk = [&]() { pool_free(s, d_fk_count); pool_free(s, d_fk_lab); pool_free(s, d_fk_bm); d_fk_count = d_fk_lab = nullptr; d_fk_bm = nullptr; };
    std::vector<Ctx> chunks;
    bool loop_live = false;
    cleanup.f = [&]() {
        if (loop_live) {
            for (Ctx &c : chunks) flush_swaps(s, c);
            if (!chunks.empty()) commit_ctx(s, chunks[0]);
            cudaDeviceSynchronize();
        }
        free_ac();
        free_fk();
        free_eq();
    };
    if (want_eq) {
        if (s->layout == PP_LAYOUT_SLAB) { return fail(PP_ERR_UNSUPPORTED, "equilibration_diagnostic is not implemented for the slab layout"); }
        ckpts.resize(80);
        ckpts.resize((size_t)pp_equil_checkpoints(cfg->n_sweeps, ckpts.data()));
        const size_t ndt = (size_t)m.D * m.T;
        CUDA_TRY(pool_alloc(s, (void **)&d_eq_sum, sizeof(double) * ndt * 2));
        CUDA_TRY(pool_alloc(s, (void **)&d_eq_snap, sizeof(double) * ndt * 2 * ckpts.size()));
        CUDA_TRY(cudaMemsetAsync(d_eq_sum, 0, sizeof(double) * ndt * 2, s->stream));
        CUDA_TRY(cudaMemsetAsync(d_eq_snap, 0, sizeof(double) * ndt * 2 * ckpts.size(), s->stream));
    }
    // Fortuin-Kasteleyn cluster updates (mod.rs:434-470): int8 layouts with unit couplings
    const bool want_oc = cfg->overlap_cluster_update_interval > 0;
    const bool want_fk = cfg->cluster_update_interval > 0;
    int64_t fk_smem_sites = 0;
    size_t fk_smem = 0;
    // multispin layout: Wolff mode only (32 clusters grow as one bit-parallel flood fill), state of one pair in shared memory
    const size_t oc_msc_smem = ((size_t)3 * m.N + (size_t)m.N * m.z) * 4 + 16;
    const bool oc_msc_ok = s->layout == PP_LAYOUT_MSC && cfg->overlap_cluster_mode == PP_CLUSTER_WOLFF && m.N <= 65536 &&
                           oc_msc_smem <= 200 * 1024;
    if (want_oc && ((s->layout != PP_LAYOUT_INT8 && !oc_msc_ok) || m.R < 2 || m.R > 64)) {
        if (s->layout != PP_LAYOUT_SLAB && m.R < 2)  // mod.rs:207-213
            return fail(PP_ERR_INVALID, "overlap cluster requires n_replicas >= max group_size (" + std::to_string(m.R) + " < 2)");
        return fail(PP_ERR_UNSUPPORTED, "overlap cluster moves (overlap_cluster_update_interval) are not implemented on the GPU sweep "
                                        "path for this handle: int8 layout, or the multispin layout with overlap_cluster_mode='wolff'");
    }
    if (want_oc && s->layout == PP_LAYOUT_MSC) {
        CUDA_TRY(cudaFuncSetAttribute(msc_houdayer_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oc_msc_smem));
        CUDA_TRY(cudaFuncSetAttribute(msc_houdayer_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oc_msc_smem));
        CUDA_TRY(cudaFuncSetAttribute(msc_houdayer_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oc_msc_smem));
        if (!s->d_nbr16) {
            const int64_t n = m.N * 2 * m.z;
            CUDA_TRY(pool_alloc(s, (void **)&s->d_nbr16, sizeof(uint16_t) * (size_t)n));
            nbr_to_u16_kernel<<<blocks_for(n, 256), 256, 0, s->stream>>>(m.nbr, s->d_nbr16, n);
            std::vector<uint16_t> site16((size_t)m.N);
            const bool permuted = s->plan.compact && !s->plan.perm.empty();
            for (int64_t i = 0; i < m.N; i++) site16[permuted ? s->plan.perm[(size_t)i] : (size_t)i] = (uint16_t)i;
            CUDA_TRY(pool_alloc(s, (void **)&s->d_site16, sizeof(uint16_t) * (size_t)m.N));
            CUDA_TRY(cudaMemcpyAsync(s->d_site16, site16.data(), sizeof(uint16_t) * (size_t)m.N, cudaMemcpyHostToDevice, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
        }
    }
    if (s->sys_comm && (want_fk || want_oc))
        return fail(PP_ERR_UNSUPPORTED, "cluster moves are not implemented for system-split handles (system_ranks > 1)");
    if (want_fk) {
        if (s->layout != PP_LAYOUT_INT8 || m.coupling_class == COUP_F32) {
            return fail(PP_ERR_UNSUPPORTED, "cluster updates (cluster_update_interval) are not implemented on the GPU sweep path for "
                                            "this handle: they need the int8 layout and couplings in {-1, 0, +1}");
        }
        std::vector<uint32_t> counts((size_t)m.T);
        for (int t = 0; t < m.T; t++) {  // fk.rs:113 with interaction = 1: u < 1 - exp(-2 / T) on the 24-bit grid
            const float p = 1.0f - expf(-2.0f * 1.0f / s->temps[(size_t)t]);
            double cnt = p > 0.0f ? std::ceil((double)p * 16777216.0) : 0.0;
            counts[(size_t)t] = (uint32_t)std::min(cnt, 16777216.0);
        }
        CUDA_TRY(pool_alloc(s, (void **)&d_fk_count, sizeof(uint32_t) * counts.size()));
        CUDA_TRY(cudaMemcpyAsync(d_fk_count, counts.data(), sizeof(uint32_t) * counts.size(), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    if (want_fk || (want_oc && s->layout == PP_LAYOUT_INT8)) {
        if ((size_t)m.N * 5 <= 200 * 1024) {  // labels (u32) + bond / activity masks (u8) in shared memory
            fk_smem_sites = m.N;
            fk_smem = (size_t)m.N * 5 + 16;
            CUDA_TRY(cudaFuncSetAttribute(fk_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fk_smem));
            CUDA_TRY(cudaFuncSetAttribute(houdayer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fk_smem));
        } else {
            CUDA_TRY(pool_alloc(s, (void **)&d_fk_lab, sizeof(uint32_t) * (size_t)(m.D * m.S * m.N)));
            CUDA_TRY(pool_alloc(s, (void **)&d_fk_bm, (size_t)(m.D * m.S * m.N)));
        }
    }
    const int64_t launches0 = s->launches;
    s->profile = s->profile_next;
    s->profile_next = false;
    s->prof_used = 0;
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));

    const bool msc = s->layout == PP_LAYOUT_MSC;
    // ---- chunk plan.  The msc3d path splits the realizations into chunks of whole word groups whose spin words fit
    // the L2 cache a few times over, runs `macro_batch` sweeps of one chunk back to back (so its words stay L2-resident
    // between the sweep, exchange and swap kernels) and spreads the chunks over a few streams, so that the
    // latency-/bandwidth-bound exchange and swap kernels of one chunk overlap the ALU-bound sweep kernel of another.
    // Realizations are independent (mod.rs:887-903 runs them on different threads), so results do not depend on this.
    int64_t macro_batch = 1;
    if (msc && s->msc3d && !s->profile && s->n_streams > 1) {
        const int64_t group_bytes = (int64_t)m.S * m.N * 4;
        int64_t gpc = s->chunk_groups;
        if (gpc <= 0 && s->chunk_bytes > 0) gpc = std::max<int64_t>(1, s->chunk_bytes / group_bytes);
        if (gpc <= 0) {
            const int64_t lo = std::max<int64_t>(1, (int64_t(8) << 20) / group_bytes), hi = std::max<int64_t>(lo, (int64_t(32) << 20) / group_bytes);
            gpc = std::min(hi, std::max(lo, (s->G + s->n_streams - 1) / s->n_streams));
        }
        const int64_t n_chunks = (s->G + gpc - 1) / gpc;
        if (n_chunks > 1) {
            pp_status stx = ensure_streams(s, (int)std::min<int64_t>(n_chunks, s->n_streams));
            if (stx != PP_OK) return stx;
            for (int64_t c = 0; c < n_chunks; c++) {
                const int64_t d0 = c * gpc * 32, D = std::min<int64_t>(m.D - d0, gpc * 32);
                chunks.push_back(chunk_ctx(s, d0, D, s->xstreams[(size_t)(c % (int64_t)s->xstreams.size())]));
            }
            for (cudaStream_t xs : s->xstreams) CUDA_TRY(cudaStreamWaitEvent(xs, s->ev0, 0));
            macro_batch = s->macro_batch;
        }
    }
    if (chunks.empty()) chunks.push_back(whole_ctx(s));
    if (s->prows) macro_batch = s->max_batch;  // plain sweeps between measurements run inside one launch
    loop_live = true;

    // packed rows in cluster mode: recorded sweeps are folded inside the sweep kernel, so a batch may run across them
    const bool cl_fold = s->prows && s->prows_cluster && !want_eq && !s->profile;
    const bool cl_batch = cl_fold && !want_ac && !want_oc && !want_fk;
    struct Step {
        uint32_t sweep_index;
        int rec_from;  // cluster mode: sweeps of the batch from this one on are recorded inside the kernel (-1: none)
        int batch;
        bool record, pt_this;
        uint32_t pt_event;
        int parity;
        int64_t sid_last;  // index, inside this call, of the batch's last sweep
    };
    std::vector<Step> steps;
    int64_t sweep_id = 0;
    // ---- small realizations: the whole per-sweep sequence runs inside rows_resident_kernel (run_rows_resident)
    if (s->rows && s->resident && !s->profile && !want_ac && !want_eq && !want_fk && !want_oc) {
        st = run_rows_resident(s, chunks[0], cfg, interrupt, on_sweep, user);
        if (st != PP_OK) return st;
        sweep_id = cfg->n_sweeps;
    }
    while (sweep_id < cfg->n_sweeps) {
        if (interrupt && *interrupt) return fail(PP_ERR_INTERRUPTED, "interrupted");  // mod.rs:406-408 (polled once per macro batch)
        // the sequence of kernel steps of this macro batch (mod.rs:405-432, 486-509, 748-796), identical for every chunk
        const int64_t mb_end = std::min<int64_t>(cfg->n_sweeps, sweep_id + macro_batch);
        steps.clear();
        uint32_t sweep_counter = s->sweep_counter, pt_event = s->pt_event_counter;
        int parity = s->next_parity;
        for (int64_t sid = sweep_id; sid < mb_end;) {
            // how many sweeps until (and including) the next one that needs a reduction or PT
            int64_t batch = 1;
            if (msc || s->prows) {  // kernels that run several sweeps per launch
                while (sid + batch - 1 < mb_end - 1) {
                    const int64_t last = sid + batch - 1;
                    const bool rec = last >= cfg->warmup_sweeps;
                    const bool ptl = cfg->pt_interval > 0 && last % cfg->pt_interval == 0;
                    const bool ocl = want_oc && last % cfg->overlap_cluster_update_interval == 0;
                    const bool fkl = want_fk && last % cfg->cluster_update_interval == 0;
                    if ((rec && !cl_batch) || ptl || ocl || fkl || want_eq || batch >= s->max_batch) break;
                    batch++;
                }
            }
            const int64_t last = sid + batch - 1;
            Step stp;
            stp.sweep_index = sweep_counter;
            stp.batch = (int)batch;
            stp.record = last >= cfg->warmup_sweeps;                                  // mod.rs:410
            stp.pt_this = cfg->pt_interval > 0 && last % cfg->pt_interval == 0;       // mod.rs:486-488
            stp.pt_event = pt_event;
            stp.parity = parity;
            stp.sid_last = last;
            stp.rec_from = -1;
            if (cl_fold && stp.record && !(want_fk && last % cfg->cluster_update_interval == 0))
                stp.rec_from = (int)std::max<int64_t>(0, std::min<int64_t>(batch - 1, cfg->warmup_sweeps - sid));
            steps.push_back(stp);
            sweep_counter += (uint32_t)batch;
            if (stp.pt_this && m.T >= 2) {
                pt_event++;
                if (cfg->pt_schedule == PP_PT_FULL_LADDER) parity = 1 - parity;       // mod.rs:793-795
            } else if (stp.pt_this) {
                pt_event++;
            }
            sid += batch;
        }
        for (Ctx &c : chunks) {
            for (const Step &stp : steps) {
                bool fused = false;
                const bool fk_this = want_fk && stp.sid_last % cfg->cluster_update_interval == 0;  // mod.rs:434-437
                const bool energy_this = stp.record || stp.pt_this || want_eq;
                st = launch_sweeps(s, c, cfg->sweep_mode, stp.sweep_index, stp.batch, cfg->exact_log, energy_this && !fk_this,
                                   stp.record, stp.record || want_eq, stp.record, &fused, stp.rec_from);
                if (st != PP_OK) return st;
                if (s->sys_comm) {  // system-split handle: what the other processes produced for their systems
                    st = sys_allgather(s, c.stream, energy_this, energy_this && stp.record, (stp.record || want_eq) && c.m.P > 0);
                    if (st != PP_OK) return st;
                }
                if (fk_this) {  // after the sweep, before the measurements (mod.rs:457-470)
                    if ((st = view_sync(s, c.stream, 1)) != PP_OK) return st;  // the cluster kernels work on the int8 view
                    fk_cluster_kernel<<<(unsigned)(c.m.D * c.m.S), FK_THREADS, fk_smem, c.stream>>>(
                        c.m, d_fk_count, stp.sweep_index + (uint32_t)stp.batch - 1u, cfg->cluster_mode == PP_CLUSTER_WOLFF ? 1 : 0,
                        fk_smem_sites, d_fk_lab, d_fk_bm);
                    s->launches++;
                    CUDA_TRY(cudaGetLastError());
                    if ((st = view_sync(s, c.stream, 0)) != PP_OK) return st;
                    if (energy_this) {
                        st = launch_energy(s, c, stp.record);
                        if (st != PP_OK) return st;
                    }
                }
                if ((stp.record || want_eq) && !fused) {
                    bool folded = false;
                    if (c.m.P > 0 || stp.record) {
                        st = launch_overlap(s, c, stp.record, &folded);                // mod.rs:527-529
                        if (st != PP_OK) return st;
                    }
                    if (stp.record && !folded) {
                        fold_kernel<<<blocks_for(c.m.D * c.m.T, 128), 128, 0, c.stream>>>(c.m, c.st, c.m.P > 0);  // mod.rs:543-578
                        s->launches++;
                    }
                }
                if (want_eq) {                                                         // mod.rs:511-541 (pre-exchange system_ids)
                    const int64_t d0 = c.m.sample_offset - s->mv.sample_offset, count_after = stp.sid_last + 1;
                    int ck = -1;
                    for (size_t i = 0; i < ckpts.size(); i++)
                        if (ckpts[i] == count_after) ck = (int)i;
                    equil_push_kernel<<<blocks_for(c.m.D * c.m.T, 128), 128, 0, c.stream>>>(
                        c.m, c.dot_link, d_eq_sum + d0 * c.m.T, d_eq_sum + (size_t)m.D * m.T + d0 * c.m.T, (long long)count_after, ck,
                        (int)ckpts.size(), d_eq_snap + (size_t)d0 * ckpts.size() * 2 * c.m.T);
                    s->launches++;
                }
                if (stp.record && want_ac) {                                           // mod.rs:580-594 (pre-exchange system_ids)
                    const int64_t d0 = c.m.sample_offset - s->mv.sample_offset;
                    auto shifted = [&](const AutocorrView &a) {
                        AutocorrView v = a;
                        if (v.ring) {
                            const int64_t o1 = d0 * c.m.T, oK = o1 * (a.K + 1);
                            v.ring += oK; v.sum_prod += oK; v.sum_o += o1; v.sum_o2 += o1;
                        }
                        return v;
                    };
                    autocorr_push_kernel<<<(unsigned)(c.m.D * c.m.T), 64, 0, c.stream>>>(c.m, c.dot_spin, shifted(ac_m), shifted(ac_q),
                                                                                         (long long)(stp.sid_last - cfg->warmup_sweeps));
                    s->launches++;
                }
                if (want_oc && stp.sid_last % cfg->overlap_cluster_update_interval == 0) {  // mod.rs:596-746
                    if (s->layout == PP_LAYOUT_MSC) {
                        st = flush_swaps(s, c);  // the words must be final (the sweep launch normally consumed the last exchange)
                        if (st != PP_OK) return st;
                        const unsigned oc_grid = (unsigned)(c.G * c.m.T * c.m.P);
                        const uint32_t oc_sweep = stp.sweep_index + (uint32_t)stp.batch - 1u;
                        if (c.m.z == 3) msc_houdayer_kernel<3><<<oc_grid, OC_THREADS, oc_msc_smem, c.stream>>>(c.m, s->d_nbr16, s->d_site16, oc_sweep, c.m.sample_offset / 32);
                        else if (c.m.z == 2) msc_houdayer_kernel<2><<<oc_grid, OC_THREADS, oc_msc_smem, c.stream>>>(c.m, s->d_nbr16, s->d_site16, oc_sweep, c.m.sample_offset / 32);
                        else msc_houdayer_kernel<0><<<oc_grid, OC_THREADS, oc_msc_smem, c.stream>>>(c.m, s->d_nbr16, s->d_site16, oc_sweep, c.m.sample_offset / 32);
                    } else {
                        if ((st = view_sync(s, c.stream, 1)) != PP_OK) return st;
                        houdayer_kernel<<<(unsigned)(c.m.D * c.m.T * c.m.P), FK_THREADS, fk_smem, c.stream>>>(
                            c.m, stp.sweep_index + (uint32_t)stp.batch - 1u, cfg->overlap_cluster_mode == PP_CLUSTER_WOLFF ? 1 : 0, fk_smem_sites,
                            d_fk_lab, d_fk_bm);
                        if ((st = view_sync(s, c.stream, 0)) != PP_OK) return st;
                    }
                    s->launches++;
                    CUDA_TRY(cudaGetLastError());
                    if (stp.pt_this) {  // mod.rs:748-756: the move changed the replicas' energies
                        st = launch_energy(s, c, false);
                        if (st != PP_OK) return st;
                    }
                }
                if (stp.pt_this) {                                                     // mod.rs:748-796
                    st = launch_pt(s, c, cfg->pt_schedule, stp.pt_event, stp.parity, true);
                    if (st != PP_OK) return st;
                }
            }
        }
        s->sweep_counter = sweep_counter;
        s->pt_event_counter = pt_event;
        s->next_parity = parity;
        if (on_sweep)
            for (int64_t sid = sweep_id; sid < mb_end; sid++) on_sweep(user, (uint64_t)sid);  // mod.rs:409
        sweep_id = mb_end;
    }
    for (Ctx &c : chunks) {  // leave the canonical state behind: no pending exchange, handle pointer on the current buffer
        st = flush_swaps(s, c);
        if (st != PP_OK) return st;
    }
    if (s->sys_comm) {  // every process ends the call with every system's configuration
        st = sys_allgather(s, s->stream, false, false, true);
        if (st != PP_OK) return st;
    }
    commit_ctx(s, chunks[0]);
    loop_live = false;
    if (chunks.size() > 1) {
        for (size_t i = 0; i < s->xstreams.size(); i++) {
            CUDA_TRY(cudaEventRecord(s->xevents[i], s->xstreams[i]));
            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->xevents[i], 0));
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    float ms = 0.0f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    double prof_ms = 0.0;
    for (size_t i = 0; i + 1 < s->prof_used; i += 2) {
        float one = 0.0f;
        if (cudaEventElapsedTime(&one, s->prof_events[i], s->prof_events[i + 1]) == cudaSuccess) prof_ms += one;
    }
    const int64_t prof_n = (int64_t)(s->prof_used / 2);
    s->profile = false;
    std::vector<double> taus;  // [2][D][T]: per-realization taus of m^2, then of q^2 (mod.rs:825-832)
    if (want_ac) {
        const int64_t ndt = m.D * m.T;
        autocorr_tau_kernel<<<blocks_for(ndt, 128), 128, 0, s->stream>>>(ndt, ac_m, (long long)std::max<int64_t>(n_rec, 0), d_tau);
        if (ac_q.ring)
            autocorr_tau_kernel<<<blocks_for(ndt, 128), 128, 0, s->stream>>>(ndt, ac_q, (long long)std::max<int64_t>(n_rec, 0), d_tau + ndt);
        taus.assign((size_t)ndt * 2, 0.0);
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        CUDA_TRY(cudaMemcpy(taus.data(), d_tau, sizeof(double) * (size_t)ndt * (ac_q.ring ? 2 : 1), cudaMemcpyDeviceToHost));
    }
    std::vector<double> eq_snap;  // [D][n_ckpt][2][T]
    if (want_eq) {
        eq_snap.assign((size_t)m.D * ckpts.size() * 2 * m.T, 0.0);
        CUDA_TRY(cudaMemcpy(eq_snap.data(), d_eq_snap, sizeof(double) * eq_snap.size(), cudaMemcpyDeviceToHost));
    }
    s->last_timing = pp_timing{ms, s->launches - launches0, prof_ms, prof_n};
    if (!out) return PP_OK;

    return collect_results(s, out, n_rec, want_eq, want_ac, ckpts, eq_snap, taus);
}

// ------------------------------------------------------------------------------------------
// state access
// slab layout: host [S][local planes][L1][L2] +-1 <-> the slabs' own planes; dir 0 = get, 1 = set
static pp_status slab_copy_spins(pp_sim *s, int8_t *host, int dir) {
    SlabState *sl = s->slab;
    const ModelView &m = s->mv;
    const int64_t per_sys = sl->local_planes() * sl->plane;
    const size_t n = (size_t)(m.S * per_sys);
    int8_t *tmp = nullptr;
    CUDA_TRY(cudaMalloc((void **)&tmp, n));
    if (dir == 1) CUDA_TRY(cudaMemcpy(tmp, host, n, cudaMemcpyHostToDevice));
    for (size_t i = 0; i < sl->parts.size(); i++) {
        const SlabView &v = sl->parts[i];
        slab_convert_kernel<<<dim3(blocks_for((int64_t)v.P * v.plane, 256), (unsigned)m.S), 256, 0, s->stream>>>(
            v, tmp, per_sys, (int64_t)i * v.P * v.plane, dir);
    }
    for (size_t i = 0; i < sl->pparts.size(); i++) {
        const SlabPView &v = sl->pparts[i];
        slabp_convert_kernel<<<dim3(blocks_for((int64_t)v.P * v.half, SLABP_THREADS), (unsigned)m.S), SLABP_THREADS, 0, s->stream>>>(
            v, tmp, per_sys, (int64_t)i * v.P * sl->plane, dir);
    }
    cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e == cudaSuccess && dir == 0) e = cudaMemcpy(host, tmp, n, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    CUDA_TRY(e);
    return PP_OK;
}

extern "C" pp_status pp_get_spins(pp_sim *s, int64_t r, int8_t *out) {
    if (!s || !out) return fail(PP_ERR_INVALID, "sim/out is NULL");
    if (r < 0 || r >= s->mv.D) return fail(PP_ERR_INVALID, "realization index out of range");
    CUDA_TRY(cudaSetDevice(s->device));
    const ModelView &m = s->mv;
    const size_t n = (size_t)m.S * m.N;
    if (s->layout == PP_LAYOUT_SLAB) return slab_copy_spins(s, out, 0);
    if (s->layout == PP_LAYOUT_MSC) {
        int8_t *tmp = nullptr;
        CUDA_TRY(pool_alloc(s, (void **)&tmp, n));
        msc_unpack_kernel<<<blocks_for((int64_t)n, 256), 256, 0, s->stream>>>(m, r, tmp);
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        CUDA_TRY(cudaMemcpy(out, tmp, n, cudaMemcpyDeviceToHost));
        pool_free(s, tmp);
    } else {
        if (s->prows || s->swords) {
            pp_status st = view_sync(s, s->stream, 1);
            if (st != PP_OK) return st;
        }
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        CUDA_TRY(cudaMemcpy(out, m.spins + (size_t)r * n, n, cudaMemcpyDeviceToHost));
    }
    return PP_OK;
}

extern "C" pp_status pp_set_spins(pp_sim *s, int64_t r, const int8_t *spins) {
    if (!s || !spins) return fail(PP_ERR_INVALID, "sim/spins is NULL");
    if (r < 0 || r >= s->mv.D) return fail(PP_ERR_INVALID, "realization index out of range");
    CUDA_TRY(cudaSetDevice(s->device));
    const ModelView &m = s->mv;
    const size_t n = (size_t)m.S * m.N;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->layout == PP_LAYOUT_SLAB) {
        pp_status st = slab_copy_spins(s, const_cast<int8_t *>(spins), 1);
        if (st != PP_OK) return st;
        st = slab_exchange(s, s->stream);
        if (st != PP_OK) return st;
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        return PP_OK;
    }
    if (s->layout == PP_LAYOUT_MSC) {
        int8_t *tmp = nullptr;
        CUDA_TRY(pool_alloc(s, (void **)&tmp, n));
        CUDA_TRY(cudaMemcpy(tmp, spins, n, cudaMemcpyHostToDevice));
        msc_pack_kernel<<<blocks_for((int64_t)n, 256), 256, 0, s->stream>>>(m, r, tmp);
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        pool_free(s, tmp);
    } else {
        if (s->prows || s->swords) {  // the int8 view of the other realizations must be current before everything is packed again
            pp_status st = view_sync(s, s->stream, 1);
            if (st != PP_OK) return st;
            CUDA_TRY(cudaStreamSynchronize(s->stream));
        }
        CUDA_TRY(cudaMemcpy(m.spins + (size_t)r * n, spins, n, cudaMemcpyHostToDevice));
        if (s->prows || s->swords) {
            pp_status st = view_sync(s, s->stream, 0);
            if (st != PP_OK) return st;
            CUDA_TRY(cudaStreamSynchronize(s->stream));
        }
    }
    return PP_OK;
}

extern "C" pp_status pp_get_system_ids(pp_sim *s, int64_t r, int64_t *out) {
    if (!s || !out) return fail(PP_ERR_INVALID, "sim/out is NULL");
    if (r < 0 || r >= s->mv.D) return fail(PP_ERR_INVALID, "realization index out of range");
    CUDA_TRY(cudaSetDevice(s->device));
    std::vector<int32_t> tmp((size_t)s->mv.S);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(tmp.data(), s->d_sid + r * s->mv.S, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); i++) out[i] = tmp[i];
    return PP_OK;
}

// With the multispin layout the words are slot-major, so re-labelling slots means moving lanes:
// unpack with the old labels, store the new labels, pack again.
extern "C" pp_status pp_set_system_ids(pp_sim *s, int64_t r, const int64_t *ids) {
    if (!s || !ids) return fail(PP_ERR_INVALID, "sim/ids is NULL");
    if (r < 0 || r >= s->mv.D) return fail(PP_ERR_INVALID, "realization index out of range");
    const ModelView &m = s->mv;
    std::vector<int32_t> tmp((size_t)m.S);
    std::vector<char> seen((size_t)m.S, 0);
    for (int k = 0; k < m.S; k++) {
        // parallel.rs:13-14: each replica's slice must stay a permutation of its own systems
        if (ids[k] < 0 || ids[k] >= m.S || seen[(size_t)ids[k]] || ids[k] / m.T != k / m.T)
            return fail(PP_ERR_INVALID, "system_ids must permute the systems of each replica ladder");
        seen[(size_t)ids[k]] = 1;
        tmp[(size_t)k] = (int32_t)ids[k];
    }
    CUDA_TRY(cudaSetDevice(s->device));
    std::vector<int8_t> spins;
    if (s->layout == PP_LAYOUT_MSC) {
        spins.resize((size_t)m.S * m.N);
        pp_status st = pp_get_spins(s, r, spins.data());
        if (st != PP_OK) return st;
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(s->d_sid + r * m.S, tmp.data(), sizeof(int32_t) * tmp.size(), cudaMemcpyHostToDevice));
    if (s->layout == PP_LAYOUT_MSC) return pp_set_spins(s, r, spins.data());
    return PP_OK;
}

extern "C" pp_status pp_get_energies(pp_sim *s, int64_t r, float *out) {
    if (!s || !out) return fail(PP_ERR_INVALID, "sim/out is NULL");
    if (r < 0 || r >= s->mv.D) return fail(PP_ERR_INVALID, "realization index out of range");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(out, s->d_energies + r * s->mv.S, sizeof(float) * (size_t)s->mv.S, cudaMemcpyDeviceToHost));
    return PP_OK;
}

// ------------------------------------------------------------------------------------------
// operator-level entry points
extern "C" pp_status pp_op_sweep(pp_sim *s, int32_t sweep_mode, uint32_t sweep_index, int32_t exact_log) {
    if (!s) return fail(PP_ERR_INVALID, "sim is NULL");
    if (sweep_mode != PP_SWEEP_METROPOLIS && sweep_mode != PP_SWEEP_GIBBS) return fail(PP_ERR_INVALID, "unknown sweep_mode");
    CUDA_TRY(cudaSetDevice(s->device));
    const bool f32 = s->mv.coupling_class == COUP_F32;
    pp_status st = ensure_tables(s, f32 && exact_log && sweep_mode == PP_SWEEP_METROPOLIS, f32 && exact_log && sweep_mode == PP_SWEEP_GIBBS);
    if (st != PP_OK) return st;
    Ctx wc = whole_ctx(s);
    st = launch_sweeps(s, wc, sweep_mode, sweep_index, 1, exact_log, false, false);
    commit_ctx(s, wc);
    if (st != PP_OK) return st;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return PP_OK;
}

extern "C" pp_status pp_op_energies_mags(pp_sim *s, float *energies, int64_t *mags) {
    if (!s) return fail(PP_ERR_INVALID, "sim is NULL");
    CUDA_TRY(cudaSetDevice(s->device));
    Ctx wc = whole_ctx(s);
    pp_status st = launch_energy(s, wc, true);
    if (st != PP_OK) return st;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const size_t n = (size_t)s->mv.D * s->mv.S;
    if (energies) CUDA_TRY(cudaMemcpy(energies, s->d_energies, sizeof(float) * n, cudaMemcpyDeviceToHost));
    if (mags) CUDA_TRY(cudaMemcpy(mags, s->d_mags, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
    return PP_OK;
}

extern "C" pp_status pp_op_overlap(pp_sim *s, int64_t *dot_spin, int64_t *dot_link) {
    if (!s) return fail(PP_ERR_INVALID, "sim is NULL");
    if (s->mv.P == 0) return fail(PP_ERR_INVALID, "overlap needs n_replicas >= 2");
    CUDA_TRY(cudaSetDevice(s->device));
    Ctx wc = whole_ctx(s);
    pp_status st = launch_overlap(s, wc);
    if (st != PP_OK) return st;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    const size_t n = (size_t)s->mv.D * s->mv.P * s->mv.T;
    if (dot_spin) CUDA_TRY(cudaMemcpy(dot_spin, s->d_dot_spin, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
    if (dot_link) CUDA_TRY(cudaMemcpy(dot_link, s->d_dot_link, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
    return PP_OK;
}

extern "C" pp_status pp_op_pt(pp_sim *s, int32_t pt_schedule, uint32_t pt_event) {
    if (!s) return fail(PP_ERR_INVALID, "sim is NULL");
    if (pt_schedule != PP_PT_SINGLE_RANDOM_EDGE && pt_schedule != PP_PT_FULL_LADDER) return fail(PP_ERR_INVALID, "unknown pt_schedule");
    CUDA_TRY(cudaSetDevice(s->device));
    pp_status st = ensure_tables(s, true, false);
    if (st != PP_OK) return st;
    Ctx wc = whole_ctx(s);
    st = launch_pt(s, wc, pt_schedule, pt_event, s->next_parity);
    if (pt_schedule == PP_PT_FULL_LADDER) s->next_parity = 1 - s->next_parity;  // mod.rs:793-795
    if (st != PP_OK) return st;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return PP_OK;
}

// ------------------------------------------------------------------------------------------
// The reference's operator signatures on HOST slices (SURVEY.md 8b "Granularity"): one realization, H2D -> kernel -> D2H around a
// temporary handle.  Unit-level parity only: a caller that sweeps repeatedly keeps a handle (pp_create / pp_sample).
struct SliceHandle {
    pp_sim *sim = nullptr;
    ~SliceHandle() { if (sim) free_sim(sim); }
};

static pp_status slice_open(const pp_model_desc *model, const int8_t *spins, const int64_t *system_ids, SliceHandle &h) {
    if (!model) return fail(PP_ERR_INVALID, "model is NULL");
    if (model->n_disorder != 1) return fail(PP_ERR_INVALID, "slice entry points take one realization (n_disorder = 1)");
    pp_status st = pp_create(model, &h.sim);
    if (st != PP_OK) return st;
    if (system_ids && (st = pp_set_system_ids(h.sim, 0, system_ids)) != PP_OK) return st;
    if (spins && (st = pp_set_spins(h.sim, 0, spins)) != PP_OK) return st;
    return PP_OK;
}

extern "C" pp_status pp_slice_sweep(const pp_model_desc *model, int32_t sweep_mode, uint32_t sweep_index, int32_t exact_log, int8_t *spins,
                                    const int64_t *system_ids) {
    if (!spins) return fail(PP_ERR_INVALID, "spins is NULL");
    SliceHandle h;
    pp_status st = slice_open(model, spins, system_ids, h);
    if (st != PP_OK) return st;
    if ((st = pp_op_sweep(h.sim, sweep_mode, sweep_index, exact_log)) != PP_OK) return st;
    return pp_get_spins(h.sim, 0, spins);
}

extern "C" pp_status pp_slice_energies_mags(const pp_model_desc *model, const int8_t *spins, float *energies, int64_t *mags) {
    if (!spins) return fail(PP_ERR_INVALID, "spins is NULL");
    SliceHandle h;
    pp_status st = slice_open(model, spins, nullptr, h);
    if (st != PP_OK) return st;
    return pp_op_energies_mags(h.sim, energies, mags);
}

extern "C" pp_status pp_slice_overlap(const pp_model_desc *model, const int8_t *spins, const int64_t *system_ids, int64_t *dot_spin,
                                      int64_t *dot_link) {
    if (!spins) return fail(PP_ERR_INVALID, "spins is NULL");
    SliceHandle h;
    pp_status st = slice_open(model, spins, system_ids, h);
    if (st != PP_OK) return st;
    return pp_op_overlap(h.sim, dot_spin, dot_link);
}

extern "C" pp_status pp_slice_pt(const pp_model_desc *model, int32_t pt_schedule, uint32_t pt_event, int32_t first_parity, const float *energies,
                                 int64_t *system_ids) {
    if (!energies || !system_ids) return fail(PP_ERR_INVALID, "energies/system_ids is NULL");
    SliceHandle h;
    pp_status st = slice_open(model, nullptr, system_ids, h);
    if (st != PP_OK) return st;
    pp_sim *s = h.sim;
    CUDA_TRY(cudaMemcpy(s->d_energies, energies, sizeof(float) * (size_t)s->mv.S, cudaMemcpyHostToDevice));
    s->next_parity = first_parity ? 1 : 0;
    if ((st = pp_op_pt(s, pt_schedule, pt_event)) != PP_OK) return st;
    return pp_get_system_ids(s, 0, system_ids);
}
