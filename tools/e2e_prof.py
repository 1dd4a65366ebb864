import time, sys, numpy as np, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import bench, peapods_b200 as pb
D=4096
Jt = torch.empty((D,)+bench.SHAPE+(3,), dtype=torch.float32, pin_memory=True); J=Jt.numpy()
bench.make_couplings(0, D, D, out=J)
temps=bench.temperatures(); seed=bench.dynamics_seed()
kw=dict(pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25, per_sample=False)
for it in range(3):
    t0=time.perf_counter()
    s=pb.IsingSimulation(list(bench.SHAPE), J, temps, 4, None, seed, layout="msc")
    torch.cuda.synchronize(); t1=time.perf_counter()
    out=s.sample(32,"metropolis",**kw); t2=time.perf_counter()
    out=s.sample(32,"metropolis",**kw); t3=time.perf_counter()
    del s; torch.cuda.synchronize(); t4=time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms, sample1 {1e3*(t2-t1):.1f} ms, sample2 {1e3*(t3-t2):.1f}, destroy {1e3*(t4-t3):.1f}")
