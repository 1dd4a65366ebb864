"""The reference's own Python package above this engine (peapods_b200/dropin.py).  Needs the reference tree, which exists in the
build container only: skipped elsewhere.  The GPU variant samples through the reference's `Ising` class."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/python")
needs_ref = pytest.mark.skipif(not REF.exists(), reason="the reference tree is mounted in the build container only")

SCRIPT = r"""
import sys
sys.path.insert(0, {root!r})
import peapods_b200.dropin as dropin
core = dropin.install({ref!r})
import peapods
from peapods import Ising, run_sweep
import peapods.spin_models as sm
import inspect
assert sm.IsingSimulation is core.IsingSimulation
assert Path(inspect.getfile(Ising)).is_relative_to({ref!r}), inspect.getfile(Ising)
import peapods.cli
print("wired")
{extra}
"""


def _run(extra=""):
    code = "from pathlib import Path\n" + SCRIPT.format(root=str(ROOT), ref=str(REF), extra=extra)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)


@needs_ref
def test_reference_package_imports_above_the_engine():
    out = _run()
    assert out.returncode == 0 and "wired" in out.stdout, out.stderr[-2000:]


@needs_ref
@pytest.mark.gpu
def test_reference_ising_class_samples_on_the_gpu():
    extra = r"""
import numpy as np
m = Ising((8, 8, 8), couplings="bimodal", temperatures=np.linspace(0.8, 1.4, 6), n_replicas=4, seed=7)
res = m.sample(200, pt_interval=1, overlap_cluster_update_interval=1, autocorrelation_max_lag=10)
assert m.sg_binder.shape == (6,) and np.all(np.isfinite(m.heat_capacity)) and "mags2_tau" in res
f = Ising((16, 16), temperatures=np.linspace(1.8, 2.8, 8), n_replicas=2, seed=1)
f.sample(300, cluster_update_interval=1, pt_interval=1)
assert f.binder_cumulant[0] > 0.6 > f.binder_cumulant[-1]
print("sampled", m._sim.layout, f._sim.layout)
"""
    out = _run(extra)
    assert out.returncode == 0 and "sampled" in out.stdout, out.stderr[-2000:]
