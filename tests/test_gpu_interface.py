"""T1: the reference's interface tests (/root/reference/tests/test_sampling_interfaces.py) for the parts that touch the
sweep path, run against `peapods_b200.Ising` — same constructor / sample keywords, result keys, shapes, dtypes, counters
and error ordering.  Cases that exercise cluster moves or the CLI are outside the path:
here they must fail BEFORE any state mutation, as the reference orders its own validation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_explicit_seed_controls_couplings_and_reset_replays_dynamics():  # test_sampling_interfaces.py:12-42
    from peapods_b200 import Ising

    temperatures = np.array([1.0, 2.0], dtype=np.float32)
    first = Ising((4, 4), couplings="bimodal", temperatures=temperatures, n_replicas=2, seed=41)
    second = Ising((4, 4), couplings="bimodal", temperatures=temperatures, n_replicas=2, seed=41)
    initial_spins = first._sim.get_spins().copy()
    np.testing.assert_array_equal(first.couplings, second.couplings)
    np.testing.assert_array_equal(initial_spins, second._sim.get_spins())
    first.sample(2, warmup_ratio=0)
    first.reset()
    np.testing.assert_array_equal(first._sim.get_spins(), initial_spins)
    first.reset(seed=99)
    seeded_reset = first._sim.get_spins().copy()
    first.reset(seed=99)
    np.testing.assert_array_equal(first._sim.get_spins(), seeded_reset)
    first.reset()
    np.testing.assert_array_equal(first._sim.get_spins(), initial_spins)


def test_same_seed_gives_the_same_results_dict():
    from peapods_b200 import Ising

    kw = dict(couplings="bimodal", temperatures=np.linspace(0.9, 1.6, 4), n_replicas=2, n_disorder=32, seed=5)
    a = Ising((4, 4, 8), **kw).sample(40, pt_interval=1)
    b = Ising((4, 4, 8), **kw).sample(40, pt_interval=1)
    for k, v in a.items():
        if k == "per_disorder":
            for kk, vv in v["parallel_tempering"].items():
                np.testing.assert_array_equal(vv, b[k]["parallel_tempering"][kk])
        elif k == "overlap_histogram":
            np.testing.assert_array_equal(np.stack(v), np.stack(b[k]))
        else:
            np.testing.assert_array_equal(v, b[k])


def test_disorder_zero_is_stable_when_disorder_count_grows():  # :45-48
    from peapods_b200 import Ising

    one = Ising((4, 4), couplings="gaussian", n_disorder=1, seed=7)
    many = Ising((4, 4), couplings="gaussian", n_disorder=3, seed=7)
    np.testing.assert_array_equal(one.couplings, many.couplings[0])


def test_result_keys_shapes_and_dtypes():  # src/lib.rs:337-412, 458-490
    from peapods_b200 import Ising

    T, R, D, N = 3, 2, 2, 16
    model = Ising((4, 4), couplings="bimodal", temperatures=np.array([1.0, 2.0, 4.0]), n_replicas=R, n_disorder=D, seed=3)
    res = model.sample(6, pt_interval=1, warmup_ratio=0)
    for k in ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4", "link_overlap",
              "link_overlap2", "link_overlap4"):
        assert res[k].shape == (T,) and res[k].dtype == np.float64, k
    assert len(res["overlap_histogram"]) == T and all(h.shape == (N + 1,) and h.dtype == np.uint64 for h in res["overlap_histogram"])
    assert res["ql_at_q_sum"].shape == (T, N + 1) and res["ql2_at_q_sum"].dtype == np.float64
    assert res["per_sample_overlap_histogram"].shape == (D, T, N + 1)
    assert res["per_sample_ql_at_q_sum"].shape == (D, T, N + 1) and res["per_sample_ql2_at_q_sum"].shape == (D, T, N + 1)
    pt = res["per_disorder"]["parallel_tempering"]
    assert pt["edge_attempts"].shape == (D, T - 1) and pt["edge_attempts"].dtype == np.uint64
    assert pt["edge_acceptances"].shape == (D, T - 1) and pt["round_trips"].shape == (D, R, T)
    # every recorded sweep puts one entry per pair into the histogram of each temperature
    assert all(int(h.sum()) == 6 * (R // 2) * D for h in res["overlap_histogram"])
    # single replica: no overlap keys (src/lib.rs:352)
    single = Ising((4, 4), temperatures=np.array([2.0, 2.5]), seed=3).sample(4, warmup_ratio=0)
    assert "overlap" not in single and "per_disorder" not in single
    assert model.binder_cumulant.shape == (T,) and model.heat_capacity.shape == (T,) and model.sg_binder.shape == (T,)


def test_full_ladder_pt_counters_accumulate_and_reset():  # :75-118 without the CMR observer
    from peapods_b200 import Ising

    model = Ising((4, 4), couplings="bimodal", temperatures=np.array([1.0, 2.0, 4.0]), n_replicas=2, seed=11)
    result = model.sample(2, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    pt = result["per_disorder"]["parallel_tempering"]
    assert pt["edge_attempts"].shape == (1, 2)
    assert np.all(pt["edge_attempts"] == 4)
    assert pt["round_trips"].shape == (1, 2, 3)
    continued = model.sample(1, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    assert np.all(continued["per_disorder"]["parallel_tempering"]["edge_attempts"] == 6)
    model.reset()
    reset = model.sample(1, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    assert np.all(reset["per_disorder"]["parallel_tempering"]["edge_attempts"] == 2)


def test_unsupported_observe_fails_before_mutation():  # :145-156
    from peapods_b200 import Ising

    model = Ising((4, 4), temperatures=np.array([2.0]), seed=13)
    before = model._sim.get_spins().copy()
    with pytest.raises(ValueError, match="requires cluster_mode='sw'"):
        model.sample(1, cluster_update_interval=1, cluster_mode="wolff", cluster_action="observe", warmup_ratio=0)
    np.testing.assert_array_equal(model._sim.get_spins(), before)


def test_invalid_autocorrelation_backend_fails_before_sampling():  # :197-206
    from peapods_b200 import Ising

    model = Ising((4, 4), temperatures=np.array([1.0, 2.0]), seed=43)
    before = model._sim.get_spins().copy()
    with pytest.raises(ValueError, match="must be 'ring' or 'fft'"):
        model.sample(4, autocorrelation_backend="other", warmup_ratio=0)
    with pytest.raises(ValueError, match="requires autocorrelation_max_lag"):
        model.sample(4, autocorrelation_backend="fft", warmup_ratio=0)
    np.testing.assert_array_equal(model._sim.get_spins(), before)


@pytest.mark.parametrize("kwargs", [
    dict(cluster_update_interval=1, cluster_action="observe"),         # FK graph observation (cluster moves themselves run)
    dict(cluster_update_interval=1, collect_cluster_stats=True),
    dict(overlap_cluster_update_interval=1, overlap_cluster_build_mode="jorg"),
    dict(overlap_cluster_update_interval=1, overlap_cluster_build_mode="cmr+houdayer"),
    dict(overlap_cluster_update_interval=1, overlap_cluster_build_mode="houd4"),
    dict(overlap_cluster_update_interval=1, overlap_cluster_action="observe"),
    dict(overlap_cluster_update_interval=1, snapshot_interval=2),
])
def test_options_outside_the_sweep_path_are_rejected_before_mutation(kwargs):
    from peapods_b200 import Ising

    model = Ising((4, 4), couplings="bimodal", temperatures=np.array([1.0, 2.0]), n_replicas=2, seed=9)
    before = model._sim.get_spins().copy()
    with pytest.raises(ValueError, match="not implemented on the GPU sweep path"):
        model.sample(3, warmup_ratio=0, **kwargs)
    np.testing.assert_array_equal(model._sim.get_spins(), before)
    model.sample(3, warmup_ratio=0)  # the handle is still usable


def test_couplings_shape_mismatch_message():  # src/lib.rs:146-149
    import peapods_b200 as pb

    with pytest.raises(ValueError, match="does not match lattice"):
        pb.IsingSimulation([4, 4], np.ones((4, 5, 2), np.float32), np.array([1.0], np.float32))


def test_interrupt_flag_raises_keyboard_interrupt():  # src/lib.rs:304-308, 327-333
    import peapods_b200 as pb

    sim = pb.IsingSimulation([8, 8], np.ones((8, 8, 2), np.float32), np.array([2.0], np.float32), 1, None, 1)
    flag = np.ones(1, dtype=np.int32)
    with pytest.raises(KeyboardInterrupt):
        sim.sample(10, "metropolis", interrupt=flag)


def test_on_sweep_callback_counts_every_sweep():  # simulation/mod.rs:409
    import peapods_b200 as pb

    sim = pb.IsingSimulation([8, 8], np.ones((8, 8, 2), np.float32), np.array([2.0, 2.5], np.float32), 1, None, 1)
    seen = []
    sim.sample(37, "metropolis", pt_interval=2, on_sweep=seen.append)
    assert seen == list(range(37))


def test_cluster_moves_on_an_automatic_multispin_model_fall_back_to_int8():
    """Round-1 review: FK updates / SW-mode Houdayer moves need cluster labels per lane, which the multispin layout does not have.
    A model built with the automatic layout is moved to int8 (configurations and system ids carried over) instead of raising;
    an explicit layout="msc" still raises."""
    import peapods_b200 as pb

    temps = np.linspace(0.9, 1.6, 4)
    model = pb.Ising((4, 4, 8), couplings="bimodal", temperatures=temps, n_replicas=2, n_disorder=32, seed=3)
    assert model._sim.layout == "msc"
    model.sample(20, pt_interval=1)
    before = [model._sim.get_spins(d).copy() for d in (0, 31)]
    ids = model._sim.get_system_ids(31).copy()
    model._fall_back_to_int8()                                                       # what sample() does on such a request
    assert model._sim.layout == "int8"
    assert all(np.array_equal(model._sim.get_spins(d), b) for d, b in zip((0, 31), before))   # configurations carried over
    assert np.array_equal(model._sim.get_system_ids(31), ids)
    fresh = pb.Ising((4, 4, 8), couplings="bimodal", temperatures=temps, n_replicas=2, n_disorder=32, seed=3)
    fresh.sample(10, cluster_update_interval=1, cluster_mode="sw", pt_interval=1)   # the automatic path: no error, int8 afterwards
    assert fresh._sim.layout == "int8" and np.all(np.isfinite(fresh.binder_cumulant))
    res = model.sample(30, pt_interval=1, overlap_cluster_update_interval=1, overlap_cluster_mode="sw")
    assert np.all(np.isfinite(res["energies"])) and model.sg_binder.shape == (4,)
    explicit = pb.Ising((4, 4, 8), couplings="bimodal", temperatures=temps, n_replicas=2, n_disorder=32, seed=3, layout="msc")
    with pytest.raises(ValueError, match="not implemented on the GPU sweep path"):
        explicit.sample(4, cluster_update_interval=1)
