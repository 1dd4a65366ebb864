import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import peapods_b200 as pb
L = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sim = pb.IsingSimulation([L, L, L], "ferro", np.asarray([4.511], np.float32), 1, None, 7, layout="slab")
sim.sample(3, "metropolis", warmup_ratio=1.0)
print("done", sim.last_sweep_loop_ms)
