"""`run_sweep`, the `.npz` writer and `peapods sweep` (python/peapods/sweep.py:100-163, 351-512, cli.py:463-566).

CPU (this container holds the reference tree and no GPU): the reference's OWN run_sweep and CLI run through peapods_b200.dropin on
a CPU stand-in with the IsingSimulation interface (tests/oracle_backend.py), this repo's mirror (peapods_b200/sweep.py) runs on
the same stand-in, and the two `.npz` files must agree key by key, bit for bit.  The key / shape / dtype list of the reference's
file is committed (tests/golden/run_sweep_npz.json); on the GPU box the mirror runs over the engine and must write that list."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/python")
GOLDEN = ROOT / "tests" / "golden" / "run_sweep_npz.json"
needs_ref = pytest.mark.skipif(not REF.exists(), reason="the reference tree is mounted in the build container only")

SIZES = [(4, 4, 4), (4, 4, 8)]
KW = dict(couplings=("bimodal",), n_replicas=2, n_disorder=3, n_sweeps=40, pt_interval=1, pt_schedule="full_ladder",
          autocorrelation_max_lag=4, equilibration_diagnostic=True, save_data=True, seed=11)
TEMPS = np.linspace(0.9, 1.8, 5)

REF_SCRIPT = r"""
import sys, json
from pathlib import Path
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
from oracle_backend import OracleIsingSimulation
import peapods_b200.dropin as dropin
dropin.install({ref!r}, backend=OracleIsingSimulation)
from peapods import run_sweep
import peapods.cli
kw = dict({kw!r})
run_sweep({sizes!r}, temperatures=np.asarray({temps!r}), output_dir={out!r}, **kw)
# the CLI: the same sweep from the command line (sys.argv), into another directory
sys.argv = ["peapods", "sweep", "--sizes", "4,4,4", "4,4,8", "--couplings", "bimodal", "--temp-min", "0.9", "--temp-max", "1.8",
            "--n-temps", "5", "--temp-scale", "linear", "--n-replicas", "2", "--n-disorder", "3", "--n-sweeps", "40", "--pt-interval", "1",
            "--pt-schedule", "full_ladder", "--autocorrelation-max-lag", "4", "--equilibration-diagnostic", "--save-data", "--seed", "11",
            "--output-dir", {out_cli!r}]
peapods.cli.main()
print("reference-ran")
"""


def _npz_signature(path):
    with np.load(path) as f:
        return {k: [list(f[k].shape), str(f[k].dtype)] for k in sorted(f.files)}


@needs_ref
def test_reference_run_sweep_and_cli_run_through_the_dropin_and_the_mirror_writes_the_same_file(tmp_path, monkeypatch):
    out_ref, out_cli, out_mine = tmp_path / "ref", tmp_path / "cli", tmp_path / "mine"
    code = REF_SCRIPT.format(root=str(ROOT), tests=str(ROOT / "tests"), ref=str(REF), kw=KW, sizes=SIZES, temps=TEMPS.tolist(),
                             out=str(out_ref), out_cli=str(out_cli))
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and "reference-ran" in run.stdout, run.stderr[-3000:]

    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_backend import OracleIsingSimulation

    import peapods_b200.spin_models as sm
    from peapods_b200.sweep import run_sweep

    monkeypatch.setattr(sm, "IsingSimulation", OracleIsingSimulation)
    res = run_sweep(SIZES, temperatures=TEMPS, output_dir=str(out_mine), **KW)
    assert list(res) == ["bimodal"] and list(res["bimodal"]) == ["4x4x4", "4x4x8"]

    ref, cli, mine = (np.load(d / "sweep_bimodal.npz") for d in (out_ref, out_cli, out_mine))
    assert sorted(ref.files) == sorted(mine.files) == sorted(cli.files)
    for k in ref.files:
        assert ref[k].dtype == mine[k].dtype and ref[k].shape == mine[k].shape, k
        assert np.array_equal(ref[k], mine[k], equal_nan=True), k        # same seeds, same backend: bit for bit
        assert np.array_equal(ref[k], cli[k], equal_nan=True), k         # the CLI is the same sweep
    # the committed signature is the reference's (regenerate with UPDATE_GOLDEN=1)
    sig = _npz_signature(out_ref / "sweep_bimodal.npz")
    import os
    if os.environ.get("UPDATE_GOLDEN"):
        GOLDEN.write_text(json.dumps(sig, sort_keys=True) + "\n")
    assert json.loads(GOLDEN.read_text()) == sig


def test_sweep_seed_derivation_and_labels_follow_the_reference():
    from peapods_b200.sweep import config_label, cumulative_overlap_ratio, run_child_seed, run_seed_words, size_label

    words = run_seed_words(11)
    assert len(words) == 4 and run_child_seed(words, "bimodal", (4, 4, 4)) != run_child_seed(words, "bimodal", (4, 4, 8))
    assert run_child_seed(words, "ferro", (8, 8)) != run_child_seed(words, "gaussian", (8, 8))
    with pytest.raises(ValueError):
        run_seed_words(-1)
    assert config_label("bimodal", "houdayer", "wolff") == "bimodal" and config_label("ferro", "jorg", "sw") == "ferro_jorg_sw"
    assert size_label((8, 8, 16)) == "8x8x16"
    hist = np.zeros((2, 1, 5), np.uint64)
    hist[0, 0, 2] = 4
    hist[1, 0, [0, 4]] = 2
    q, ratio, mean, median = cumulative_overlap_ratio(hist)
    assert q.tolist() == [0.0, 0.5, 1.0] and mean[0].tolist() == [0.5, 0.5, 1.0] and ratio[0, 2] == 1.0


@pytest.mark.gpu
def test_run_sweep_over_the_engine_writes_the_reference_npz_layout(tmp_path):
    from peapods_b200.sweep import run_sweep

    res = run_sweep(SIZES, temperatures=TEMPS, output_dir=str(tmp_path), **KW)
    model = res["bimodal"]["4x4x8"]
    assert np.all(np.isfinite(model.sg_binder)) and model.per_disorder["parallel_tempering"]["round_trips"].shape == (3, 2, 5)
    assert _npz_signature(tmp_path / "sweep_bimodal.npz") == json.loads(GOLDEN.read_text())
