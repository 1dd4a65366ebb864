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
    if "per_disorder" in first:
        pt = {}
        for name in ("edge_attempts", "edge_acceptances", "round_trips"):
            pt[name] = np.concatenate([p["result"]["per_disorder"]["parallel_tempering"][name] for p in parts], axis=0)
        out["per_disorder"] = {"parallel_tempering": pt}
    return out


def gather_merge(result, per_sample_means, n_replicas, group=None, dst=0):
    """Collective: gather every rank's part on ``dst`` and merge (returns None elsewhere).  Uses the object
    collectives of ``torch.distributed`` (a few MB of scalars; works on the NCCL and the gloo backend)."""
    import torch.distributed as dist

    part = {"result": result, "per_sample_means": np.asarray(per_sample_means), "n_replicas": int(n_replicas)}
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
        if self.count == 0:
            raise ValueError("more ranks than 32-realization groups: nothing to do on this rank")
        shard = couplings(self.first, self.count) if callable(couplings) else np.asarray(couplings)[self.first:self.first + self.count]
        self.n_replicas = 1 if n_replicas is None else int(n_replicas)
        self.sim = IsingSimulation(lattice_shape, shard, temperatures, n_replicas, neighbor_offsets, seed, layout=layout,
                                   device=self.rank if device is None else device, sample_offset=self.first)

    def sample(self, *args, **kwargs):
        """Every rank samples its block; rank 0 returns the merged dict, the others None."""
        local = self.sim.sample(*args, **kwargs)
        return gather_merge(local, self.sim.last_per_sample_means, self.n_replicas)
