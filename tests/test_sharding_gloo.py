"""N > 1 host logic on CPU: two gloo ranks each run their block of realizations (the oracle stands in for the GPU
engine here -- no CUDA in this container), gather_merge() on rank 0 must rebuild the unsharded run."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent

SHAPE, TEMPS, R, D = (4, 4, 4), np.linspace(0.8, 1.6, 4).astype(np.float32), 2, 96


def _couplings(n=D):
    rng = np.random.default_rng(5)
    return (2 * rng.integers(0, 2, size=(n,) + SHAPE + (3,)) - 1).astype(np.float32)


def _run_block(first, count, n=D):
    import oracle
    import peapods_b200  # noqa: F401  (the package must import without a GPU)
    from peapods_b200 import colouring

    colour, _ = colouring(SHAPE)
    sim = oracle.Sim(SHAPE, _couplings(n)[first:first + count], TEMPS, n_replicas=R, seed=77, rng_mode=oracle.RNG_PHILOX_MSC,
                     colour=colour, sample_offset=first)
    res = sim.sample(30, "metropolis", pt_interval=1, pt_schedule="full_ladder", autocorrelation_max_lag=4)
    res["overlap_histogram"] = [h for h in res["overlap_histogram"]]
    return res, sim.last_per_sample_means, sim.last_per_sample_taus


def _worker(rank, world, port, out_path, n=D):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from peapods_b200.sharded import gather_merge, shard_bounds

    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_bounds(n, world, rank)
    if count == 0:  # more ranks than 32-realization groups: the rank still takes part in the collective
        merged = gather_merge(None, None, R)
    else:
        res, means, taus = _run_block(first, count, n)
        merged = gather_merge(res, means, R, per_sample_taus=taus)
    if rank == 0:
        np.savez(out_path, **{k: np.asarray(v) for k, v in merged.items() if k != "per_disorder"},
                 **{"pt_" + k: v for k, v in merged["per_disorder"]["parallel_tempering"].items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything_in_groups_of_32():
    sys.path.insert(0, str(ROOT))
    from peapods_b200.sharded import shard_bounds

    for n, world in ((4096, 8), (96, 2), (100, 3), (33, 4), (32, 2)):
        blocks = [shard_bounds(n, world, r) for r in range(world)]
        assert sum(c for _, c in blocks) == n
        pos = 0
        for first, count in blocks:
            assert first == pos or count == 0
            assert first % 32 == 0 or count == 0
            pos += count


def test_two_gloo_ranks_reproduce_the_unsharded_run(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    merged = np.load(out_path)
    sys.path.insert(0, str(ROOT))
    ref, _, _ = _run_block(0, D)
    for k in ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4", "link_overlap",
              "link_overlap2", "link_overlap4", "mags2_tau", "overlap2_tau"):
        assert np.array_equal(merged[k], ref[k]), k            # ordered sum over realizations: bit for bit
    assert np.array_equal(merged["overlap_histogram"], np.stack(ref["overlap_histogram"]))
    for k in ("ql_at_q_sum", "ql2_at_q_sum"):
        np.testing.assert_allclose(merged[k], ref[k], rtol=1e-12, atol=0)   # partial sums re-associated
    for k in ("per_sample_overlap_histogram", "per_sample_ql_at_q_sum"):
        assert np.array_equal(merged[k], ref[k]), k
    for k, v in ref["per_disorder"]["parallel_tempering"].items():
        assert np.array_equal(merged["pt_" + k], v), k


def test_an_empty_rank_and_a_partial_word_group_still_merge_to_the_unsharded_run(tmp_path):
    """40 realizations over 3 ranks: 32 + 8 + 0 (the empty rank used to raise while the others waited in the
    gather, and the 8-realization shard must keep the batch's multispin keys)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "merged40.npz")
    mp.spawn(_worker, args=(3, port, out_path, 40), nprocs=3, join=True)
    merged = np.load(out_path)
    sys.path.insert(0, str(ROOT))
    ref, _, _ = _run_block(0, 40, 40)
    for k in ("mags", "mags2", "energies", "energies2", "overlap2", "overlap4", "link_overlap", "mags2_tau"):
        assert np.array_equal(merged[k], ref[k]), k
    assert np.array_equal(merged["overlap_histogram"], np.stack(ref["overlap_histogram"]))
    assert np.array_equal(merged["per_sample_overlap_histogram"], ref["per_sample_overlap_histogram"])
    for k, v in ref["per_disorder"]["parallel_tempering"].items():
        assert np.array_equal(merged["pt_" + k], v), k


# ---- slab decomposition of one lattice: host-side plan and NCCL-token plumbing (the halo exchange itself is NCCL on
# GPUs: tests/test_gpu_parity.py emulates the ranks on one device, tools/slab_check.py runs real ranks) ----
def test_slab_plan_cuts_even_slabs():
    import pytest

    sys.path.insert(0, str(ROOT))
    from peapods_b200.sharded import slab_plan

    assert [slab_plan(1024, 8, r) for r in range(8)] == [(128 * r, 128) for r in range(8)]
    assert slab_plan(8, 2, 1) == (4, 4)
    with pytest.raises(ValueError):
        slab_plan(12, 4, 0)   # 3 planes per slab: odd, the colour of a site would depend on the cut


def test_system_plan_cuts_contiguous_blocks_of_systems():
    import pytest

    sys.path.insert(0, str(ROOT))
    from peapods_b200.sharded import system_plan

    assert [system_plan(128, 8, r) for r in range(8)] == [(16 * r, 16) for r in range(8)]   # BASELINE configs[2]: 64 temps x 2 replicas
    assert system_plan(6, 2, 1) == (3, 3)
    with pytest.raises(ValueError):
        system_plan(6, 4, 0)


def _token_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from peapods_b200.sharded import broadcast_token

    dist.init_process_group("gloo", rank=rank, world_size=world)
    token = broadcast_token(lambda: bytes(range(128)), None)
    Path(out_dir, f"token{rank}.bin").write_bytes(token)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_share_the_bootstrap_token(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_token_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "token0.bin").read_bytes() == (tmp_path / "token1.bin").read_bytes() == bytes(range(128))
