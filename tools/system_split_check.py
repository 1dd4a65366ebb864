"""Real multi-GPU check of the system-split handles (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/system_split_check.py

The S systems of one realization are split over the ranks; energies / magnetisations / configurations travel with ncclAllGather
inside the engine.  Every rank must hold the spins, the system ids and the result dict of the CPU oracle's (unsplit) run, bit
for bit: triangular ferromagnet with Gibbs sweeps (the shape of BASELINE configs[2]) and a 3-D +-J model with parallel tempering
and replica overlaps."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import peapods_b200 as pb  # noqa: E402
from peapods_b200.sharded import SystemSplitIsingSimulation  # noqa: E402

TRI = [[1, 0], [0, 1], [1, -1]]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import oracle

    ok = True
    tc = 4.0 / np.log(3.0)
    rng = np.random.default_rng(3)
    cases = [
        ((16, 16), TRI, "ferro", np.linspace(tc - 0.4, tc + 0.4, 2 * world), 2, "gibbs", dict(pt_interval=None)),
        ((16, 16), TRI, "ferro", np.linspace(tc - 0.4, tc + 0.4, 2 * world), 2, "gibbs", dict(pt_interval=1, pt_schedule="full_ladder")),
        ((4, 4, 8), None, (2 * rng.integers(0, 2, size=(4, 4, 8, 3)) - 1).astype(np.float32), np.linspace(0.8, 1.6, world), 4, "metropolis",
         dict(pt_interval=1, pt_schedule="single_random_edge", autocorrelation_max_lag=3)),
    ]
    for shape, offsets, coup, temps, R, mode, kw in cases:
        temps = np.asarray(temps, np.float32)
        sim = SystemSplitIsingSimulation(shape, coup, temps, R, offsets, 99)
        z = len(offsets) if offsets else len(shape)
        J = np.ones(tuple(shape) + (z,), np.float32) if isinstance(coup, str) else coup
        colour, _ = pb.colouring(shape, offsets)
        cpu = oracle.Sim(shape, J, temps, n_replicas=R, offsets=offsets, seed=99, rng_mode=oracle.RNG_PHILOX, colour=colour)
        for n_sweeps in (3, 21):
            rg = sim.sample(n_sweeps, mode, warmup_ratio=0.25, **kw)
            rc = cpu.sample(n_sweeps, mode, warmup_ratio=0.25, **kw)
            for k, v in rc.items():
                if k == "per_disorder":
                    same = all(np.array_equal(v["parallel_tempering"][f], rg[k]["parallel_tempering"][f]) for f in v["parallel_tempering"])
                elif k == "overlap_histogram":
                    same = np.array_equal(np.stack(rg[k]), np.asarray(v))
                else:
                    same = np.array_equal(np.asarray(rg[k]), np.asarray(v))
                if not same:
                    ok = False
                    print(f"rank {rank}: MISMATCH {shape} {mode} {k}")
            if not np.array_equal(sim.get_spins(), cpu.spins(0)):
                ok = False
                print(f"rank {rank}: spins differ {shape} {mode}")
        if rank == 0:
            print(f"system_split_check shape={shape} R={R} T={len(temps)} {mode} {kw}: {'ok' if ok else 'DIFFERS'}", flush=True)
        del sim
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("SYSTEM_SPLIT_CHECK", "PASS" if int(flag.item()) else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
