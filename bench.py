#!/usr/bin/env python
"""bench.py — spin-flip attempts/ns of the sweep path on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): 3-D Edwards-Anderson +-J, L=16, 32 temperatures 0.8-1.4,
4 replicas, 4096 disorder samples per GPU, Metropolis + parallel tempering every sweep + overlap,
couplings from the reference's own generator (seed 42, python/peapods/spin_models.py:104-126),
warmup_ratio 0.25 (3/4 of the sweeps pay the energy / magnetisation / overlap reductions).

A "step" is one `sample(n_sweeps=SWEEPS_PER_STEP)` call.  Three measurements:
  value   device-resident: a persistent engine handle, K `pp_sample` calls, device time of the sweep
          loops (CUDA events on the launch stream), max over ranks.
  e2e     through the reference-facing API with host buffers: every step constructs
          `IsingSimulation` from the pinned host coupling array (H2D), samples, and reads the result
          dict back (D2H); wall clock bracketed by synchronize, max over ranks.
  roofline  the sweep kernel alone: algorithmic bytes per launch / its mean launch duration,
          measured live with CUDA events around every sweep-kernel launch (separate, untimed pass).
The same JSON line carries sub-records measured in the same invocation (each with its own small, fixed step count):
  c2_strong   BASELINE configs[1] at its NAMED size (4096 samples in total) split over the N ranks (strong scaling)
  c5          BASELINE configs[4] (1024^3 ferromagnet at T_c) slab-decomposed over the N ranks: value, e2e, roofline, halo bytes,
              and an N-rank parity check (NCCL ranks == the same slabs kept on one device, bit for bit)
  c3_split    (N > 1) BASELINE configs[2] with its 128 systems split over the N ranks (ncclAllGather of energies / configurations),
              with a parity check against the unsplit run
  c1, c3, c4  (N = 1) BASELINE configs[0], [2], [3] through the public API
  e2e_default_api   the headline's e2e with the reference's default return set (per-realization histograms included)
`--impl reference` times the CPU restatement of the reference's rayon path (oracle, typewriter order +
xoshiro streams, threads over realizations) on the box's host cores, on a bounded sample of the same
workload.  The Rust crate itself cannot be built in this image (no cargo/rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "spin_flip_attempts_per_ns"
UNIT = "attempts/ns"
SHAPE = (16, 16, 16)
N_TEMPS, T_LO, T_HI, N_REPLICAS = 32, 0.8, 1.4, 4
SAMPLES_PER_GPU = 4096
Z = 3
SEED = 42
B_ALG_MSC = 0.125 + 0.125 + (Z / 8.0) / (N_TEMPS * N_REPLICAS)  # bytes per attempt
# dram__bytes_read.sum + dram__bytes_write.sum of one msc3d_kernel launch covering the FULL 4096 samples (recorded sweep), mean of the
# two launches of the `ncu --set full` capture summarised in profiles/r2_summary.md (350.6 MB read + 273.6 MB written); a launch
# over another sample count is scaled by it
NCU_DRAM_BYTES_PER_4096_SAMPLES = 624.2e6


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def temperatures():
    return np.linspace(T_LO, T_HI, N_TEMPS).astype(np.float32)


def make_couplings(first_sample: int, n_samples: int, total: int, out: np.ndarray | None = None) -> np.ndarray:
    """The reference generator (spin_models.py:107-126): realization r draws from child r of the coupling
    SeedSequence, so a shard of samples equals the same slice of the unsharded run."""
    from peapods_b200.spin_models import seed_material

    coupling_seq, _ = seed_material(SEED)
    children = coupling_seq.spawn(total)[first_sample:first_sample + n_samples]
    single = SHAPE + (Z,)
    if out is None:
        out = np.empty((n_samples,) + single, dtype=np.float32)
    for i, child in enumerate(children):
        out[i] = 2 * np.random.default_rng(child).integers(0, 2, size=single) - 1
    return out


def dynamics_seed() -> int:
    from peapods_b200.spin_models import seed_material

    return seed_material(SEED)[1]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thread.join(timeout=2)

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_run(n_samples: int, n_sweeps: int, n_threads: int, repeats: int = 1):
    """Reference-faithful CPU path (oracle: typewriter order, xoshiro per system, +-J lookup, energies +
    overlap recomputed per recorded sweep, PT label swaps; threads over realizations like
    simulation/mod.rs:887-903).  Returns (attempts/ns, seconds)."""
    import oracle

    temps = temperatures()
    J = make_couplings(0, n_samples, SAMPLES_PER_GPU)
    sim = oracle.Sim(SHAPE, J, temps, n_replicas=N_REPLICAS, seed=dynamics_seed(), rng_mode=oracle.RNG_XOSHIRO)
    attempts = float(np.prod(SHAPE)) * N_TEMPS * N_REPLICAS * n_samples * n_sweeps
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        sim.sample(n_sweeps, "metropolis", pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25,
                   n_threads=n_threads, per_sample=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return attempts / (best * 1e9), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_samples = max(cores, 16) * 2
    n_sweeps = min(args.sweeps_per_step, 256)  # bounded CPU step (the metric is a rate: attempts per ns)
    # bounded: scale the sample so one step stays near a few seconds
    for _ in range(args.warmup):
        cpu_baseline_run(n_samples, max(2, n_sweeps // 8), cores)
    times, vals = [], []
    for _ in range(args.steps):
        v, dt = cpu_baseline_run(n_samples, n_sweeps, cores)
        vals.append(v)
        times.append(dt)
    attempts = float(np.prod(SHAPE)) * N_TEMPS * N_REPLICAS * n_samples * n_sweeps
    value = attempts * len(times) / (sum(times) * 1e9)
    sample = (f"{n_samples} of {SAMPLES_PER_GPU} disorder samples x {n_sweeps} sweeps per step, "
              f"same lattice/temperatures/replicas/PT/overlap")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 multispin words / int8 spins (CPU: int8 + f32)",
        "data": "synthetic", "config": dict(workload_config(args, n_samples), sweeps_per_step=n_sweeps),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference CPU path (oracle/pp_oracle.c); the Rust crate cannot be built here",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, samples_per_gpu):
    return {
        "workload": "C2: 3-D EA +-J L=16, 32 temps 0.8-1.4, 4 replicas, Metropolis + PT(single_random_edge, every sweep) + overlap",
        "lattice": list(SHAPE), "n_temps": N_TEMPS, "n_replicas": N_REPLICAS, "samples_per_gpu": samples_per_gpu,
        "sweeps_per_step": args.sweeps_per_step, "warmup_ratio": 0.25, "pt_interval": 1, "layout": "msc (32 samples / u32 word)",
        "sharding": "disorder samples split across GPUs, no data-path collective",
        "l2": "working set (256 MiB spin words + per-sample histograms) exceeds the 126 MB L2; no flush needed",
    }


# ---------------------------------------------------------------------------------------------------------------------
# sub-records of the one JSON line
TRI_OFFSETS = [[1, 0], [0, 1], [1, -1]]


class Ranks:
    """barrier / reductions over the ranks of the (already initialised) default process group"""

    def __init__(self, world, rank, local_rank):
        self.world, self.rank, self.local_rank = world, rank, local_rank

    def barrier(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def _reduce(self, x, op):
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        import torch.distributed as dist

        return self._reduce(x, dist.ReduceOp.MAX)

    def sum(self, x):
        import torch.distributed as dist

        return self._reduce(x, dist.ReduceOp.SUM)


def slab_rank_parity(ranks: Ranks):
    """The N-rank slab path (ncclSend / ncclRecv halos, ncclAllReduce of the bond counts) against the SAME decomposition kept on
    one device (device copies instead of NCCL; that form is checked bit for bit against the CPU restatement by tests/): spins of
    every rank, energies and the statistics must be identical.  Byte storage and bit-packed storage, PT over three temperatures."""
    import torch
    import torch.distributed as dist

    import peapods_b200 as pb
    from peapods_b200.sharded import SlabIsingSimulation

    world, rank = ranks.world, ranks.rank
    ok = True
    for shape in ((8 * world, 6, 16), (4 * world, 16, 128)):
        temps = np.asarray([4.0, 4.511, 5.0], np.float32)
        kw = dict(warmup_ratio=0.25, pt_interval=1, pt_schedule="full_ladder")
        sim = SlabIsingSimulation(shape, temps, 99)
        res = sim.sample(25, "metropolis", **kw)
        mine = torch.from_numpy(sim.get_spins().astype(np.int8)).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        if world > 1:
            dist.gather(mine, parts, dst=0)
        else:
            parts = [mine]
        if rank == 0:
            one = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 99, layout="slab", slab_ranks=world, slab_rank=-1,
                                     device=ranks.local_rank)
            ref = one.sample(25, "metropolis", **kw)
            T, per = len(temps), int(np.prod(shape)) // world
            full = np.concatenate([p.cpu().numpy().reshape(T, per) for p in parts], axis=1).reshape(-1)
            ok = ok and np.array_equal(full, one.get_spins(0))
            ok = ok and all(np.array_equal(res[k], ref[k]) for k in ("mags", "mags2", "mags4", "energies", "energies2"))
            del one
        del sim
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.broadcast(flag, src=0)
    return "PASS" if int(flag.item()) else "FAIL"


def bench_c5(ranks: Ranks, extent: int, n_sweeps: int, steps: int, warmup: int, clock_index=None):
    """BASELINE configs[4]: one extent^3 ferromagnet at T_c cut into one slab per rank; returns the sub-record (every rank)."""
    from peapods_b200.sharded import SlabIsingSimulation

    world = ranks.world
    shape = (extent, extent, extent)
    temps = np.asarray([4.511], dtype=np.float32)  # T_c of the 3-D Ising model (tests/utils.py:9 in the reference)
    attempts_step = float(extent) ** 3 * n_sweeps
    kw = dict(warmup_ratio=0.25)
    sim = SlabIsingSimulation(shape, temps, dynamics_seed())
    packed = bool(sim.sim.slab_packed)
    for _ in range(warmup):
        sim.sample(n_sweeps, "metropolis", **kw)
    ranks.barrier()
    dev_ms, launches = 0.0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.sample(n_sweeps, "metropolis", **kw)
        dev_ms += sim.sim.last_sweep_loop_ms
        launches += sim.sim.last_kernel_launches
    ranks.barrier()
    wall_ms = ranks.max(1e3 * (time.perf_counter() - t0))
    dev_ms = ranks.max(dev_ms)
    value = attempts_step * steps / (dev_ms * 1e6)
    res = sim.sample(n_sweeps, "metropolis", profile=True, **kw)
    k_ms = ranks.max(sim.sim.last_sweep_kernel_ms)
    peak, peak_src = peaks()
    b_alg = 0.25 if packed else 2.0  # one bit (byte) read + one written per attempt
    alg = b_alg * attempts_step / world
    achieved = alg / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    energy = float(res["energies"][0])
    del sim

    def e2e_step():
        s = SlabIsingSimulation(shape, temps, dynamics_seed())
        out = s.sample(n_sweeps, "metropolis", **kw)
        nb = sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray))
        del s
        return nb

    for _ in range(max(warmup, 1)):
        e2e_step()
    ranks.barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(steps):
        d2h = e2e_step()
    ranks.barrier()
    e2e_ms = ranks.max(1e3 * (time.perf_counter() - t0))
    half_plane = extent * extent // (16 if packed else 1)  # bytes of one colour of one plane (packed) / of one plane (bytes)
    return {
        "workload": f"C5: single 3-D Ising ferromagnet {extent}^3 at T_c=4.511, checkerboard Metropolis, slab-decomposed along x0 over "
                    f"{world} GPU(s)" + (" with NCCL halo exchange" if world > 1 else ""),
        "value": value, "unit": UNIT, "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup, "sweeps_per_step": n_sweeps,
        "ms_per_step": dev_ms / steps, "wall_ms_per_step": wall_ms / steps,
        "e2e": {"value": attempts_step * steps / (e2e_ms * 1e6), "unit": UNIT, "h2d_bytes_per_step": int(temps.nbytes),
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps},
        "layout": "slab, one bit per spin (32 same-colour sites of a row per u32 word)" if packed else "slab (u8, stride geometry)",
        "dtype": "u32 bit-sliced (packed draws, integer acceptance table)" if packed else "u8",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "alg_bytes_per_attempt": b_alg, "kernel": "slabp_sweep_kernel" if packed else "slab_sweep_kernel",
                     "sweep_kernel_ms": k_ms, "peak_source": peak_src,
                     "note": "the kernel is bound by the generator's multiplies and three compares per site, not by HBM (DESIGN.md)"},
        "halo_bytes_per_half_step_per_rank": 0 if world == 1 else 2 * half_plane,
        "gpu_launches": launches, "energy_per_spin_last_step": energy,
        "l2": f"{extent ** 3 / world / (8 if packed else 1) / 2 ** 20:.0f} MiB of spins per GPU vs 126 MB L2",
    }


def bench_c3_split(ranks: Ranks):
    """BASELINE configs[2] (triangular 256^2, Gibbs, 64 temps, 2 replicas = 128 systems of 64 Ki sites) with the SYSTEMS split over the
    N ranks: energies / magnetisations per measurement and configurations per recorded sweep travel with
    ncclAllGather inside the engine.  Parity: the split run equals the unsplit run on one GPU (rank 0), result dict and spins."""
    import torch
    import torch.distributed as dist

    import peapods_b200 as pb
    from peapods_b200.sharded import SystemSplitIsingSimulation

    world, rank = ranks.world, ranks.rank
    tc = 4.0 / np.log(3.0)
    temps = np.linspace(tc - 0.4, tc + 0.4, 64).astype(np.float32)
    shape, R, n_sweeps = (256, 256), 2, 400
    seed = dynamics_seed()
    sim = SystemSplitIsingSimulation(shape, "ferro", temps, R, TRI_OFFSETS, seed)
    sim.sample(100, "gibbs")
    ranks.barrier()
    best = None
    for _ in range(2):
        sim.sample(n_sweeps, "gibbs")
        ms = ranks.max(sim.sim.last_sweep_loop_ms)
        best = ms if best is None else min(best, ms)
    value = float(np.prod(shape)) * len(temps) * R * n_sweeps / (best * 1e6)
    # parity on the same lattice, with exchange events: every rank's dict and spins against the unsplit handle on rank 0
    kw = dict(warmup_ratio=0.25, pt_interval=1, pt_schedule="full_ladder")
    sim.reset()
    res = sim.sample(24, "gibbs", **kw)
    spins = sim.get_spins()
    ok = True
    if rank == 0:
        one = pb.IsingSimulation(list(shape), "ferro", temps, R, TRI_OFFSETS, seed, layout="int8", device=ranks.local_rank)
        ref = one.sample(24, "gibbs", **kw)
        ok = np.array_equal(spins, one.get_spins(0))
        for k, v in ref.items():
            if k == "per_disorder":
                ok = ok and all(np.array_equal(v["parallel_tempering"][f], res[k]["parallel_tempering"][f]) for f in v["parallel_tempering"])
            elif k == "overlap_histogram":
                ok = ok and np.array_equal(np.stack(v), np.stack(res[k]))
            else:
                ok = ok and np.array_equal(np.asarray(v), np.asarray(res[k]))
        del one
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.broadcast(flag, src=0)
    peak, _ = peaks()
    del sim
    return {"workload": f"C3: 2-D triangular ferromagnet 256x256, Gibbs, 64 temps around 4/ln3, 2 replicas; the 128 systems split over {world} GPU(s)",
            "value": value, "unit": UNIT, "scaling": "strong", "n_gpus": world, "sweeps": n_sweeps, "ms": best,
            "alg_bytes_per_attempt": 0.25, "hbm_roofline_frac": value * 0.25 / (world * peak),
            "layout": "int8 API, one bit per spin in the engine (packed rows)",
            "collectives_per_recorded_sweep": 0 if world == 1 else 3,
            "rank_parity": "PASS" if int(flag.item()) else "FAIL",
            "rank_parity_note": "split run == unsplit run on one GPU, bit for bit (spins, statistics, exchange counters), 24 Gibbs sweeps with "
                                "full-ladder exchanges"}


def bench_small_configs():
    """BASELINE configs[0], [2], [3] through the public API on this GPU (N = 1 only): device time of the sweep loop, best of 2."""
    import peapods_b200 as pb

    peak, _ = peaks()
    out = {}

    def run(key, label, model, n_sweeps, mode, b_alg, **kw):
        model.sample(max(4, n_sweeps // 4), mode, **kw)
        best, launches = None, 0
        for _ in range(2):
            model.sample(n_sweeps, mode, **kw)
            dev = model._sim.last_sweep_loop_ms
            best, launches = (dev, model._sim.last_kernel_launches) if best is None or dev < best else (best, launches)
        attempts = float(model.n_spins) * model.n_temps * model.n_replicas * model.n_disorder * n_sweeps
        v = attempts / best / 1e6
        sim = model._sim
        note = ""
        if sim.rows_packed:
            note, b_alg = " (one bit per spin: packed rows, system resident in shared memory)", 0.25
        elif sim.sys_words:  # fp32 couplings: 1 bit read + 1 bit written per attempt, 4 z' bytes of couplings per site shared by S systems
            note = " (one bit per spin: the same site of 32 systems of a realization per word, couplings read once per 32 systems)"
            b_alg = 0.25 + 4.0 * model.n_neighbors / (model.n_temps * model.n_replicas)
        out[key] = {"workload": label, "value": v, "unit": UNIT, "sweeps": n_sweeps, "ms": best, "gpu_launches": launches,
                    "layout": sim.layout + note, "alg_bytes_per_attempt": b_alg, "hbm_roofline_frac": v * b_alg / peak}

    m = pb.Ising((32, 32), "ferro", np.linspace(1.5, 3.0, 16), n_replicas=2, seed=SEED)
    run("c1", "C1: 2-D Ising ferromagnet 32x32, Metropolis + PT every sweep, 16 temps 1.5-3.0, 2 replicas, 5000 sweeps (README quickstart)",
        m, 5000, "metropolis", 2.0, pt_interval=1)
    tc = 4.0 / np.log(3.0)
    m = pb.Ising((256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), n_replicas=2, neighbor_offsets=TRI_OFFSETS, seed=SEED)
    run("c3", "C3: 2-D triangular ferromagnet 256x256 (custom neighbor_offsets), Gibbs, 64 temps around 4/ln3, 2 replicas",
        m, 400, "gibbs", 2.0)
    m = pb.Ising((32, 32, 32), "gaussian", np.linspace(0.8, 1.8, 48), n_replicas=4, n_disorder=512, seed=SEED)
    run("c4", "C4: 3-D EA Gaussian 32^3, 48 temps 0.8-1.8, 4 replicas, 512 disorder samples, Metropolis + PT + overlap",
        m, 20, "metropolis", 2.0 + 4.0 * 3 / 192, pt_interval=1, per_sample=False)
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    import peapods_b200 as pb
    from peapods_b200 import _lib

    _lib.load()  # fails loudly if the CUDA extension is missing: there is no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise RuntimeError("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    D = args.samples_per_gpu
    temps = temperatures()
    first = rank * D
    # pinned host coupling buffer (the e2e leg copies from it every step)
    Jt = torch.empty((D,) + SHAPE + (Z,), dtype=torch.float32, pin_memory=True)
    J = Jt.numpy()
    make_couplings(first, D, D * world, out=J)
    seed = dynamics_seed()
    n_sweeps = args.sweeps_per_step
    attempts_step = float(np.prod(SHAPE)) * N_TEMPS * N_REPLICAS * D * n_sweeps
    kw = dict(pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25, per_sample=False)

    # ---- device-resident leg -------------------------------------------------------------------
    sim = pb.IsingSimulation(list(SHAPE), J, temps, N_REPLICAS, None, seed, layout="msc", device=local_rank,
                             sample_offset=first)
    assert sim.layout == "msc"
    for _ in range(args.warmup):
        sim.sample(n_sweeps, "metropolis", **kw)
    barrier()
    dev_ms, launches = 0.0, 0
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sim.sample(n_sweeps, "metropolis", **kw)
            dev_ms += sim.last_sweep_loop_ms
            launches += sim.last_kernel_launches
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    total_attempts = sum_over_ranks(attempts_step * args.steps)
    value = total_attempts / (dev_ms * 1e6)
    clock_summary = clocks.summary()

    # ---- roofline pass (untimed): CUDA events around every sweep-kernel launch ------------------
    res = sim.sample(n_sweeps, "metropolis", profile=True, **kw)
    del res
    k_ms, k_n = sim.last_sweep_kernel_ms, max(sim.last_sweep_kernel_launches, 1)
    k_n_expected = n_sweeps  # profile mode: one launch per sweep over all samples
    peak, peak_src = peaks()
    alg_bytes_per_launch = B_ALG_MSC * attempts_step / k_n
    achieved = alg_bytes_per_launch / (k_ms / k_n * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": args.traffic if args.traffic is not None else NCU_DRAM_BYTES_PER_4096_SAMPLES * D / 4096.0 * (k_n_expected / k_n),
                "traffic_source": "ncu --set full at D = 4096 in one launch (profiles/r2_summary.md), scaled by the sample count of this launch",
                "kernel": "msc3d_kernel (sweep + energy / magnetisation / overlap / fold)", "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes_per_launch, "launches_timed": k_n, "kernel_ms_mean": k_ms / k_n,
                "kernel_share_of_step": k_ms / max(sim.last_sweep_loop_ms, 1e-9)}
    # ---- context figure (untimed leg, same handle): sweeps alone -- no PT, nothing recorded, 16 sweeps per launch (macro_batch; PP_MACRO_BATCH raises it up to 64).  This is
    # the regime in which the sweep kernel approaches the HBM roofline; the headline above pays PT + reductions every sweep.
    kw_pure = dict(pt_interval=None, warmup_ratio=1.0, per_sample=False)
    sim.sample(64, "metropolis", **kw_pure)
    sim.sample(256, "metropolis", **kw_pure)
    pure_ms = max_over_ranks(sim.last_sweep_loop_ms)
    pure_value = sum_over_ranks(float(np.prod(SHAPE)) * N_TEMPS * N_REPLICAS * D * 256) / (pure_ms * 1e6)
    sweeps_only = {"value": pure_value, "unit": UNIT, "hbm_roofline_frac": pure_value * B_ALG_MSC / (world * peak),
                   "note": "256 sweeps without parallel tempering or recording (16 sweeps per launch), device time"}
    del sim

    # ---- end-to-end leg: host buffers in, result dict out, every step ----------------------------
    def e2e_step():
        s = pb.IsingSimulation(list(SHAPE), J, temps, N_REPLICAS, None, seed, layout="msc", device=local_rank,
                               sample_offset=first)
        out = s.sample(n_sweeps, "metropolis", **kw)
        nb = sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray))
        nb += sum(h.nbytes for h in out["overlap_histogram"])
        nb += sum(v.nbytes for v in out["per_disorder"]["parallel_tempering"].values())
        nb += s.last_per_sample_means.nbytes
        del s
        return nb

    for _ in range(max(args.warmup, 1)):  # same W warm-up steps as the device leg (the stream-ordered pool settles)
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        d2h = e2e_step()
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    e2e_value = total_attempts / (e2e_ms * 1e6)
    h2d = J.nbytes + temps.nbytes

    # ---- the end-of-run reduction across ranks (untimed): ordered sum over realizations on rank 0 ----------------
    if world > 1:
        from peapods_b200.sharded import gather_merge

        s = pb.IsingSimulation(list(SHAPE), J, temps, N_REPLICAS, None, seed, layout="msc", device=local_rank,
                               sample_offset=first)
        merged = gather_merge(s.sample(8, "metropolis", **kw), s.last_per_sample_means, N_REPLICAS)
        if rank == 0:
            assert np.all(np.isfinite(merged["energies"])) and len(merged["overlap_histogram"]) == N_TEMPS
        del s

    # ---- sub-records of the same line (module docstring) ---------------------------------------------
    ranks = Ranks(world, rank, local_rank)
    sub = {}
    if not args.no_sub_records:
        # BASELINE configs[1] at its named size, split over the ranks (strong scaling; N = 1 is the headline itself)
        if world > 1:
            Ds = SAMPLES_PER_GPU // world
            Js = make_couplings(rank * Ds, Ds, SAMPLES_PER_GPU)
            ssim = pb.IsingSimulation(list(SHAPE), Js, temps, N_REPLICAS, None, seed, layout="msc", device=local_rank,
                                      sample_offset=rank * Ds)
            for _ in range(2):
                ssim.sample(n_sweeps, "metropolis", **kw)
            barrier()
            s_ms = 0.0
            for _ in range(3):
                ssim.sample(n_sweeps, "metropolis", **kw)
                s_ms += ssim.last_sweep_loop_ms
            barrier()
            s_ms = max_over_ranks(s_ms)
            s_val = float(np.prod(SHAPE)) * N_TEMPS * N_REPLICAS * SAMPLES_PER_GPU * n_sweeps * 3 / (s_ms * 1e6)
            sub["c2_strong"] = {"workload": f"C2 at its named size: {SAMPLES_PER_GPU} disorder samples in total, {Ds} per GPU", "value": s_val,
                                "unit": UNIT, "scaling": "strong", "n_gpus": world, "steps": 3, "warmup": 2, "ms_per_step": s_ms / 3,
                                "hbm_roofline_frac": s_val * B_ALG_MSC / (world * peak)}
            del ssim
        else:
            sub["c2_strong"] = {"workload": f"C2 at its named size: {SAMPLES_PER_GPU} disorder samples in total", "value": value, "unit": UNIT,
                                "scaling": "strong", "n_gpus": 1, "note": "N = 1: the headline itself"}
        sub["c5"] = bench_c5(ranks, args.c5_extent, 32, 3, 2)
        sub["c5"]["rank_parity"] = slab_rank_parity(ranks)
        sub["c5"]["rank_parity_note"] = ("N NCCL ranks == the same N slabs on one device (device copies), bit for bit: spins, energies, "
                                         "statistics; byte and bit-packed storage, PT over three temperatures")
        if world > 1:
            sub["c3_split"] = bench_c3_split(ranks)
        if world == 1:
            sub.update(bench_small_configs())
            # the reference API's default return set: per-realization histograms [D][T][N+1] x 3 (src/lib.rs:385-411), 12.9 GB at C2
            kw_def = dict(kw, per_sample=True)

            def e2e_default():
                s = pb.IsingSimulation(list(SHAPE), J, temps, N_REPLICAS, None, seed, layout="msc", device=local_rank, sample_offset=first)
                out = s.sample(n_sweeps, "metropolis", **kw_def)
                nb = sum(v.nbytes for v in out.values() if isinstance(v, np.ndarray))
                del s, out
                return nb

            e2e_default()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nb = 0
            for _ in range(2):
                nb = e2e_default()
            torch.cuda.synchronize()
            def_ms = 1e3 * (time.perf_counter() - t0) / 2
            sub["e2e_default_api"] = {"value": attempts_step / (def_ms * 1e6), "unit": UNIT, "ms_per_step": def_ms, "steps": 2, "warmup": 1,
                                      "d2h_bytes_per_step": nb, "h2d_bytes_per_step": h2d,
                                      "note": "per_sample=True: what the reference returns by default when D > 1 and R >= 2; the headline "
                                              "e2e above runs with per_sample=False (aggregated histograms only)"}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns, nsw = 4 * max(cores, 16), 12
        v, dt = cpu_baseline_run(ns, nsw, cores)
        if dt < 6.0:  # aim for 10-30 s of CPU work
            nsw = int(min(600, max(nsw, nsw * 15.0 / max(dt, 1e-3))))
            v, dt = cpu_baseline_run(ns, nsw, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} of {D} disorder samples x {nsw} sweeps ({dt:.1f} s), same lattice/temps/replicas/PT/overlap"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (multispin words, 32 samples/word; integer acceptance table)", "data": "synthetic",
            "config": workload_config(args, D),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps, "clocks": clock_summary,
            "roofline": roofline, "hbm_roofline_frac_whole_step": value * B_ALG_MSC / (world * peak),
            "sweeps_only": sweeps_only, "cpu_baseline": cpu,
        }
        line["e2e"]["note"] = "per_sample=False (aggregated histograms); the default return set is timed in e2e_default_api"
        line.update(sub)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# --workload c5: BASELINE configs[4] alone, one 3-D ferromagnet at T_c cut into slabs over the ranks (strong scaling)
def run_c5(args):
    import torch
    import torch.distributed as dist

    from peapods_b200 import _lib

    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ranks = Ranks(world, rank, local_rank)
    with ClockSampler(local_rank) as clocks:
        rec = bench_c5(ranks, args.c5_extent, args.sweeps_per_step, args.steps, args.warmup)
    rec["rank_parity"] = slab_rank_parity(ranks)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        Lc, nsw = 128, 4  # the reference runs one realization with one system on ONE core (mod.rs:874-885, parallel.rs:39)
        temps = np.asarray([4.511], dtype=np.float32)
        o = oracle.Sim((Lc, Lc, Lc), np.ones((Lc, Lc, Lc, 3), np.float32), temps, n_replicas=1, seed=dynamics_seed(),
                       rng_mode=oracle.RNG_XOSHIRO)
        t0 = time.perf_counter()
        o.sample(nsw, "metropolis", warmup_ratio=0.25, n_threads=1)
        dt = time.perf_counter() - t0
        cpu = {"value": float(Lc) ** 3 * nsw / (dt * 1e9), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{Lc}^3 lattice x {nsw} sweeps ({dt:.1f} s): one system = one thread in the reference"}
    if rank == 0:
        L = args.c5_extent
        line = {
            "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": rec["dtype"], "data": "synthetic",
            "config": {"workload": rec["workload"], "lattice": [L, L, L], "n_temps": 1, "n_replicas": 1,
                       "sweeps_per_step": args.sweeps_per_step, "warmup_ratio": 0.25, "layout": rec["layout"], "l2": rec["l2"]},
            "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"], "wall_ms_per_step": rec["wall_ms_per_step"],
            "clocks": clocks.summary(), "roofline": rec["roofline"], "cpu_baseline": cpu,
            "halo_bytes_per_half_step_per_rank": rec["halo_bytes_per_half_step_per_rank"], "rank_parity": rec["rank_parity"],
            "energy_per_spin_last_step": rec["energy_per_spin_last_step"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--sweeps-per-step", type=int, default=None,
                    help="sweeps of one sample() call = one step (default: 1024 for c2, 32 for c5; the CPU arm caps its steps at 256)")
    ap.add_argument("--samples-per-gpu", type=int, default=SAMPLES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="only the headline (skip c2_strong / c5 / c1 / c3 / c4 / e2e_default_api)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from an ncu --set full capture")
    ap.add_argument("--workload", choices=("c2", "c5"), default="c2", help="c2: BASELINE headline (default); c5: one large lattice in slabs")
    ap.add_argument("--c5-extent", type=int, default=1024)
    args = ap.parse_args()
    # stdout carries exactly one JSON line: native libraries that write to file descriptor 1 (NCCL prints its version
    # there when NCCL_DEBUG is set) are sent to stderr, Python's sys.stdout keeps the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.warmup < 3:
        args.warmup = max(args.warmup, 0)
    if args.workload == "c5" and args.impl == "ours":
        if args.sweeps_per_step is None:
            args.sweeps_per_step = 32
        return run_c5(args)
    if args.sweeps_per_step is None:
        args.sweeps_per_step = 1024
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
