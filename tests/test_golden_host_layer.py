"""The host layer above the drop-in boundary against golden vectors produced by the REFERENCE's own
python/peapods/spin_models.py (tools/make_golden.py, run where /root/reference is mounted): seed derivation, coupling
generation, the argument lists crossing the boundary (`IsingSimulation(...)`, `.sample(...)`, `.reset(...)`) and the
post-processed observables must be identical.  CPU only: the engine class is replaced by the recording stub the
generator used for `peapods._core`."""
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def gen():
    spec = importlib.util.spec_from_file_location("make_golden", ROOT / "tools" / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def replay(gen):
    import peapods_b200.spin_models as sm

    recorder = gen.Recorder()
    saved = sm.IsingSimulation
    sm.IsingSimulation = recorder.make_core()
    try:
        yield gen.run_cases(sm.Ising, recorder)
    finally:
        sm.IsingSimulation = saved


def test_dynamics_seeds_match_the_reference(gen):
    from peapods_b200.spin_models import seed_material

    gold = np.load(GOLD / "host_layer.npz")
    ours = [seed_material(int(s))[1] for s in gold["seed_list"]]
    assert np.array_equal(np.asarray(ours, dtype=np.uint64), gold["dynamics_seeds"])


def test_couplings_and_observables_match_the_reference(gen, replay):
    gold = np.load(GOLD / "host_layer.npz")
    names = {k.split("/")[0] for k in gold.files if "/" in k}
    assert names == {name for name, _ in gen.CASES}
    for name, (arrays, _) in replay.items():
        keys = {k.split("/", 1)[1] for k in gold.files if k.startswith(name + "/")}
        assert keys == set(arrays), name
        for k, v in arrays.items():
            g = gold[f"{name}/{k}"]
            assert g.dtype == v.dtype and np.array_equal(g, v), f"{name}/{k}"


def test_calls_crossing_the_boundary_match_the_reference(gen, replay):
    """Engine-only keywords aside (layout / device, which the reference's constructor does not have), the stub sees the
    reference's argument lists: same names, same values, same None-ing of unused cluster options."""
    gold = json.loads((GOLD / "host_layer.json").read_text())
    for name, (_, log) in replay.items():
        ours = json.loads(json.dumps(log))
        assert len(ours) == len(gold[name]), name
        for (kind_o, args_o), (kind_g, args_g) in zip(ours, gold[name]):
            assert kind_o == kind_g
            extra = set(args_o) - set(args_g)
            assert extra <= {"layout", "device"}, (name, kind_o, extra)
            assert {k: v for k, v in args_o.items() if k not in extra} == args_g, (name, kind_o)
