"""Sharding the batch of disorder realizations over the GPUs of one box (one process per GPU).

The reference runs realizations on different rayon threads and then aggregates
(/root/reference/spin-sim/src/simulation/mod.rs:887-939, statistics/results.rs:165-180, 250-259,
statistics/overlap.rs:106-152).  Here every rank owns a contiguous block of realizations (a multiple of 32 for
the multispin layout, with its global ``sample_offset`` so that seeds do not depend on the sharding); there is no
data-path collective.  Only the end-of-run reduction crosses ranks: every rank contributes its per-realization
means ``[D_rank, 11, T]`` and its summed histograms, and rank 0 rebuilds exactly what the reference's ordered sum
over realizations gives (per-temperature means bit for bit; integer histograms bit for bit; the f64
``ql_at_q_sum`` arrays are sums of per-rank partial sums, i.e. equal up to f64 re-association)."""
from __future__ import annotations

import numpy as np

MEAN_KEYS = ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4",
             "link_overlap", "link_overlap2", "link_overlap4")


def shard_bounds(n_disorder: int, world: int, rank: int, multiple: int = 32):
    """(first, count) of the realizations rank ``rank`` owns: contiguous blocks of whole ``multiple``-sized groups,
    earlier ranks take the remainder groups."""
    groups = -(-n_disorder // multiple)
    base, extra = divmod(groups, world)
    g0 = rank * base + min(rank, extra)
    g1 = g0 + base + (1 if rank < extra else 0)
    first = min(g0 * multiple, n_disorder)
    last = min(g1 * multiple, n_disorder)
    return first, last - first


def merge_results(parts):
    """parts: per rank, in rank order, a dict with ``result`` (the rank's ``sample()`` dict), ``per_sample_means``
    ``[D_rank, 11, T]`` and ``n_replicas``.  Returns the dict the unsharded run returns."""
    parts = [p for p in parts if p is not None and p["per_sample_means"].shape[0] > 0]
    means = np.concatenate([p["per_sample_means"] for p in parts], axis=0)
    D, _, T = means.shape
    has_overlap = parts[0]["n_replicas"] >= 2
    out = {}
    for k, name in enumerate(MEAN_KEYS):
        if k >= 5 and not has_overlap:
            break
        acc = np.zeros(T, dtype=np.float64)
        for d in range(D):  # results.rs:165-180: realization order
            acc += means[d, k]
        out[name] = acc / float(D)
    first = parts[0]["result"]
    if has_overlap:
        hist = None
        for p in parts:
            h = np.stack(p["result"]["overlap_histogram"])
            hist = h.copy() if hist is None else hist + h
        out["overlap_histogram"] = [hist[t].copy() for t in range(T)]
        for name in ("ql_at_q_sum", "ql2_at_q_sum"):
            acc = np.zeros_like(first[name])
            for p in parts:
                acc += p["result"][name]
            out[name] = acc
        for name in ("per_sample_overlap_histogram", "per_sample_ql_at_q_sum", "per_sample_ql2_at_q_sum"):
            if all(name in p["result"] for p in parts) and D > 1:
                out[name] = np.concatenate([p["result"][name] for p in parts], axis=0)
    if all(p.get("per_sample_taus") is not None for p in parts):  # results.rs:217-231, 269-274: ordered mean of the taus
        taus = np.concatenate([p["per_sample_taus"] for p in parts], axis=0)  # [D, 2, T]
        for k, name in enumerate(("mags2_tau", "overlap2_tau")):
            if name not in first:
                continue
            acc = np.zeros(T, dtype=np.float64)
            for d in range(D):
                acc += taus[d, k]
            out[name] = acc / float(D)
    if all(p.get("per_sample_equil") is not None for p in parts) and "equil_sweeps" in first:  # results.rs:231-247, 275-282
        eq = np.concatenate([p["per_sample_equil"] for p in parts], axis=0)  # [D, n_ckpt, 2, T]
        out["equil_sweeps"] = first["equil_sweeps"]
        for k, name in enumerate(("equil_energy_avg", "equil_link_overlap_avg")):
            acc = np.zeros(eq.shape[1:2] + (T,), dtype=np.float64)
            for d in range(D):
                acc += eq[d, :, k]
            out[name] = acc / float(D)
    if "per_disorder" in first:
        pt = {}
        for name in ("edge_attempts", "edge_acceptances", "round_trips"):
            pt[name] = np.concatenate([p["result"]["per_disorder"]["parallel_tempering"][name] for p in parts], axis=0)
        out["per_disorder"] = {"parallel_tempering": pt}
    return out


def gather_merge(result, per_sample_means, n_replicas, group=None, dst=0, per_sample_taus=None, per_sample_equil=None):
    """Collective: gather every rank's part on ``dst`` and merge (returns None elsewhere).  Uses the object
    collectives of ``torch.distributed`` (a few MB of scalars; works on the NCCL and the gloo backend)."""
    import torch.distributed as dist

    part = None if result is None else {  # None: a rank without realizations (it still takes part in the collective)
        "result": result, "per_sample_means": np.asarray(per_sample_means), "n_replicas": int(n_replicas),
        "per_sample_taus": None if per_sample_taus is None else np.asarray(per_sample_taus),
        "per_sample_equil": None if per_sample_equil is None else np.asarray(per_sample_equil)}
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_results([part])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(part, gathered, dst=dst, group=group)
    return merge_results(gathered) if rank == dst else None


class ShardedIsingSimulation:
    """``IsingSimulation`` over this rank's block of realizations (src/lib.rs:106-174 semantics for the whole batch:
    pass the FULL coupling array or a callable ``first, count -> couplings`` that generates the shard)."""

    def __init__(self, lattice_shape, couplings, n_disorder, temperatures, n_replicas=None, neighbor_offsets=None, seed=None,
                 *, layout="auto", device=None, rank=None, world=None):
        import torch.distributed as dist

        from ._core import IsingSimulation

        live = dist.is_available() and dist.is_initialized()
        self.rank = (dist.get_rank() if live else 0) if rank is None else rank
        self.world = (dist.get_world_size() if live else 1) if world is None else world
        self.n_disorder = int(n_disorder)
        self.first, self.count = shard_bounds(self.n_disorder, self.world, self.rank)
        self.n_replicas = 1 if n_replicas is None else int(n_replicas)
        self.sim = None
        if self.count == 0:  # more ranks than 32-realization groups: this rank only takes part in the gather
            return
        shard = couplings(self.first, self.count) if callable(couplings) else np.asarray(couplings)[self.first:self.first + self.count]
        dev = self.rank if device is None else device
        if layout == "auto" and self.n_disorder >= 32 and not isinstance(shard, str) and np.any(np.asarray(shard) < 0):
            # The layout is a property of the WHOLE batch (trajectories depend on it: word-group keys and shared draws in the
            # multispin layout, per-realization keys in int8): a last shard of fewer than 32 realizations must not fall back
            # to int8 on its own.  Partial word groups are supported by the multispin kernels.
            try:
                self.sim = IsingSimulation(lattice_shape, shard, temperatures, n_replicas, neighbor_offsets, seed, layout="msc",
                                           device=dev, sample_offset=self.first)
            except ValueError:  # not a +-1 model (or too many directions): what the unsharded handle would decide too
                self.sim = None
            layout = "int8"
        if self.sim is None:
            self.sim = IsingSimulation(lattice_shape, shard, temperatures, n_replicas, neighbor_offsets, seed, layout=layout,
                                       device=dev, sample_offset=self.first)

    def sample(self, *args, **kwargs):
        """Every rank samples its block; rank 0 returns the merged dict, the others None."""
        if self.sim is None:
            return gather_merge(None, None, self.n_replicas)
        local = self.sim.sample(*args, **kwargs)
        return gather_merge(local, self.sim.last_per_sample_means, self.n_replicas,
                            per_sample_taus=getattr(self.sim, "last_per_sample_taus", None),
                            per_sample_equil=getattr(self.sim, "last_per_sample_equil", None))


def slab_plan(extent0: int, world: int, rank: int):
    """(first plane, planes) of rank ``rank`` when a lattice of ``extent0`` planes is cut into ``world`` slabs
    (pp_slab.cuh: equal slabs with an even number of planes each, so the checkerboard colour survives the cut)."""
    if extent0 % (2 * world) != 0:
        raise ValueError(f"shape[0]={extent0} must be a multiple of 2 * {world} ranks")
    planes = extent0 // world
    return rank * planes, planes


def broadcast_token(make_token, group=None, n_bytes=128):
    """Rank 0 calls ``make_token()`` (-> ``n_bytes`` bytes); every rank returns the same bytes (NCCL or gloo group)."""
    import torch
    import torch.distributed as dist

    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    buf = torch.zeros(n_bytes, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == 0:
        buf.copy_(torch.frombuffer(bytearray(make_token()), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


class SlabIsingSimulation:
    """ONE 3-D ferromagnet cut along dimension 0 over the ranks of the initialized ``torch.distributed`` group (one
    process per GPU; BASELINE config 5).  Rank 0 creates the NCCL bootstrap token, ``torch.distributed`` broadcasts
    its 128 bytes (plumbing only), and the engine itself exchanges halo planes with ncclSend / ncclRecv, overlapped with
    the interior update (pp_slab.cuh).  Every rank replays the same parallel-tempering decisions from the all-reduced
    integer energies, so ``sample()`` returns the same dict on every rank; ``get_spins()`` returns this rank's planes."""

    def __init__(self, lattice_shape, temperatures, seed=None, *, device=None, group=None):
        import torch
        import torch.distributed as dist

        from ._core import IsingSimulation, nccl_unique_id

        live = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if live else 0
        self.world = dist.get_world_size(group) if live else 1
        self.first_plane, self.planes = slab_plan(int(lattice_shape[0]), self.world, self.rank)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        # The engine keeps one communicator per (device, world, rank) for the life of the process, so only the first handle of a
        # world uses the token.  It is still broadcast for every handle: the collective keeps the ranks' handle construction in
        # step (measured at N = 2, 1024^3: 15 ms per construct + sample step with it, 58 ms without — ranks that drift apart
        # serialise on each other's first halo exchange).
        token = broadcast_token(nccl_unique_id, group) if self.world > 1 else None
        self.sim = IsingSimulation(list(lattice_shape), "ferro", temperatures, 1, None, seed, layout="slab", device=device,
                                   slab_ranks=self.world, slab_rank=self.rank if self.world > 1 else 0, nccl_unique_id=token)

    def sample(self, *args, **kwargs):
        return self.sim.sample(*args, **kwargs)

    def get_spins(self):
        return self.sim.get_spins(0)

    def reset(self, seed=None):
        self.sim.reset(seed)


def system_plan(n_systems: int, world: int, rank: int):
    """(first system, systems) of rank ``rank`` when the S = n_replicas * n_temps systems of one realization are split over
    ``world`` processes in contiguous blocks (pp_model_desc.system_ranks; the reference parallelises over systems the same
    way, spin-sim/src/parallel.rs:36-40)."""
    if n_systems % world != 0:
        raise ValueError(f"n_replicas * n_temps = {n_systems} must be a multiple of {world} ranks")
    per = n_systems // world
    return rank * per, per


class SystemSplitIsingSimulation:
    """ONE realization of many large systems (BASELINE config 3: 128 systems of 64 Ki sites) split by SYSTEM over the ranks of the
    initialized ``torch.distributed`` group (one process per GPU).  Every rank sweeps its own block of systems; the engine
    all-gathers energies / magnetisations per measurement or exchange event and configurations per recorded sweep (NCCL), and
    every rank replays the same exchange decisions and statistics, so ``sample()`` returns the same dict on every rank — the dict
    the unsplit run returns, bit for bit — and ``get_spins()`` the whole realization."""

    def __init__(self, lattice_shape, couplings, temperatures, n_replicas=None, neighbor_offsets=None, seed=None, *, device=None,
                 group=None):
        import torch
        import torch.distributed as dist

        from ._core import IsingSimulation, nccl_unique_id

        live = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if live else 0
        self.world = dist.get_world_size(group) if live else 1
        R = 1 if n_replicas is None else int(n_replicas)
        self.first_system, self.n_systems = system_plan(R * len(np.asarray(temperatures).reshape(-1)), self.world, self.rank)
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        token = broadcast_token(nccl_unique_id, group) if self.world > 1 else None
        self.sim = IsingSimulation(list(lattice_shape), couplings, temperatures, n_replicas, neighbor_offsets, seed, layout="int8",
                                   device=device, system_ranks=self.world, system_rank=self.rank, nccl_unique_id=token)

    def sample(self, *args, **kwargs):
        return self.sim.sample(*args, **kwargs)

    def get_spins(self):
        return self.sim.get_spins(0)

    def reset(self, seed=None):
        self.sim.reset(seed)
