"""Real multi-GPU check of the slab-decomposed lattice (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py

Every rank holds L0/N planes; halos travel with ncclSend/ncclRecv inside the engine.  Rank 0 gathers the slabs and
compares spins, energies and the result dict bit for bit with the CPU oracle run on the whole lattice."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import peapods_b200 as pb  # noqa: E402
from peapods_b200.sharded import SlabIsingSimulation  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ok = True
    for shape, temps in (((8 * world, 6, 16), [4.0, 4.511, 5.0]), ((4 * world, 16, 32), [4.511]),
                         ((8 * world, 6, 64), [4.0, 4.511, 5.0]), ((2 * world, 16, 128), [4.511])):  # the last two: one bit per spin
        temps = np.asarray(temps, np.float32)
        sim = SlabIsingSimulation(shape, temps, 99)
        results = []
        for n_sweeps, interval in ((3, None), (25, 1)):
            results.append(sim.sample(n_sweeps, "metropolis", warmup_ratio=0.25, pt_interval=interval, pt_schedule="full_ladder"))
        mine = torch.from_numpy(sim.get_spins().astype(np.int8)).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            import oracle

            T = len(temps)
            per = int(np.prod(shape)) // world
            full = np.concatenate([p.cpu().numpy().reshape(T, per) for p in parts], axis=1).reshape(-1)
            colour, _ = pb.colouring(shape)
            cpu = oracle.Sim(shape, np.ones(tuple(shape) + (3,), np.float32), temps, n_replicas=1, seed=99, colour=colour,
                             rng_mode=oracle.RNG_PHILOX_PACKED if sim.sim.slab_packed else oracle.RNG_PHILOX)
            for (n_sweeps, interval), rg in zip(((3, None), (25, 1)), results):
                rc = cpu.sample(n_sweeps, "metropolis", warmup_ratio=0.25, pt_interval=interval, pt_schedule="full_ladder")
                for k in ("mags", "mags2", "mags4", "energies", "energies2"):
                    if not np.array_equal(rg[k], rc[k]):
                        ok = False
                        print(f"MISMATCH {shape} {k}", rg[k], rc[k])
            same = np.array_equal(full, cpu.spins(0))
            ok = ok and same
            print(f"slab_check shape={shape} ranks={world}: spins {'ok' if same else 'DIFFER'}, energies {results[-1]['energies']}")
        del sim
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("SLAB_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
