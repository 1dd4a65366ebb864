"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every
symbol include/peapods_b200.h declares; host-only helpers agree with the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def pb():
    import __graft_entry__ as entry

    entry.build()
    import peapods_b200

    return peapods_b200


def test_library_exports_every_declared_symbol(pb):
    import ctypes

    from peapods_b200 import _lib

    header = (ROOT / "include" / "peapods_b200.h").read_text()
    declared = set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", header))
    declared -= {"pp_sim"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.pp_abi_version() == 4


def test_no_cpu_fallback_without_a_gpu(pb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pb.IsingSimulation([4, 4], np.ones((4, 4, 2), np.float32), np.array([1.0], np.float32))


def test_product_code_never_touches_the_oracle():
    for path in (ROOT / "peapods_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h", ".cpp"}:
            text = path.read_text()
            assert "oracle" not in text.lower() or path.name == "build.py", f"{path} mentions the oracle"


@pytest.mark.parametrize("shape,offsets,expected", [
    ((8, 8), None, 2), ((16, 16, 16), None, 2), ((256, 256), [[1, 0], [0, 1], [1, -1]], 4),
    ((6, 6), [[1, 0], [0, 1], [1, -1]], 3), ((2, 2), None, 2), ((4, 4, 4), [[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1]], 2),
    ((3, 3), None, 3),
])
def test_colouring_is_proper_and_minimal_where_known(pb, oracle, shape, offsets, expected):
    colour, n = pb.colouring(shape, offsets)
    assert n == expected and colour.max() == n - 1
    assert oracle.Lattice(shape, offsets).colouring_is_valid(colour)


@pytest.mark.parametrize("shape,offsets", [((5, 7), None), ((3, 5, 4), None), ((5, 5), [[1, 0], [0, 1], [1, -1]]),
                                           ((4, 4, 4), [[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1]])])
def test_colouring_fallbacks_are_proper(pb, oracle, shape, offsets):
    colour, n = pb.colouring(shape, offsets)
    assert oracle.Lattice(shape, offsets).colouring_is_valid(colour)
    assert n <= 2 * (len(shape) if offsets is None else len(offsets)) + 1


def test_self_neighbour_lattices_are_rejected(pb):
    with pytest.raises(ValueError, match="onto itself"):
        pb.colouring((1, 4))


@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
def test_host_lookup_tables_equal_the_oracle(pb, oracle, mode):
    temps = np.array([0.5, 0.8, 1.1, 2.269, 4.511, 9.0], np.float32)
    for z in (2, 3, 6):
        mine = pb.metropolis_lookup(temps, z, mode)
        fn = oracle.lib().orc_metropolis_accepted_count if mode == "metropolis" else oracle.lib().orc_gibbs_accepted_count
        ref = np.array([[fn(float(t), ec) for ec in range(-2 * z, 2 * z + 1)] for t in temps], np.uint32)
        assert np.array_equal(mine, ref)
    with pytest.raises(ValueError):
        pb.metropolis_lookup([0.0], 2, mode)


def test_seed_material_and_couplings_follow_the_reference_recipe(pb):
    from peapods_b200.spin_models import make_couplings, seed_material

    seq, dyn = seed_material(41)
    seq2, dyn2 = seed_material(41)
    assert dyn == dyn2 and dyn != seed_material(42)[1]
    one = make_couplings("gaussian", (4, 4), 2, 1, seed_material(7)[0])
    many = make_couplings("gaussian", (4, 4), 2, 3, seed_material(7)[0])
    assert np.array_equal(one, many[0])  # tests/test_sampling_interfaces.py:45-48
    b = make_couplings("bimodal", (4, 4), 2, 2, seq)
    assert set(np.unique(b)) == {-1.0, 1.0} and b.dtype == np.float32 and b.shape == (2, 4, 4, 2)
    with pytest.raises(ValueError, match="non-negative"):
        seed_material(-1)
    assert pb._lib.load().pp_realization_seed(5, 3) == __import__("oracle").lib().orc_realization_seed(5, 3)
