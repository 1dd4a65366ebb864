"""T0: the oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §4 / §8c).  Each test cites the reference test it re-expresses."""
import numpy as np
import pytest

F24 = 1 << 24


# ---------------------------------------------------------------- RNG primitives
def test_splitmix64_first_output(oracle):
    # published first output of SplitMix64 seeded with 0 (Vigna's reference code)
    assert oracle.lib().orc_splitmix64(0) == 0xE220A8397B1DCDAF


def test_xoshiro256starstar_published_vector(oracle):
    # authors' vector, state {1,2,3,4} (also rand_xoshiro's own unit test for this generator)
    s = np.array([1, 2, 3, 4], dtype=np.uint64)
    got = [oracle.lib().orc_xoshiro_next_u64(s.ctypes.data) for _ in range(10)]
    assert got == [
        11520, 0, 1509978240, 1215971899390074240, 1216172134540287360,
        607988272756665600, 16172922978634559625, 8476171486693032832,
        10595114339597558777, 2904607092377533576,
    ]


def test_philox4x32_10_random123_kats(oracle):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert oracle.philox4x32_10([f, f, f, f], [f, f]).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox4x32_7_random123_kats(oracle):
    # Random123 kat_vectors: philox4x32 7 rounds -- the RNG-SPEC v2 generator
    f = 0xFFFFFFFF
    for rounds in (7, None):
        assert oracle.philox4x32([0, 0, 0, 0], [0, 0], rounds).tolist() == [0x5F6FB709, 0x0D893F64, 0x4F121F81, 0x4F730A48]
        assert oracle.philox4x32([f, f, f, f], [f, f], rounds).tolist() == [0x5207DDC2, 0x45165E59, 0x4D8EE751, 0x8C52F662]
        assert oracle.philox4x32([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], rounds).tolist() == [
            0x4DFCCABA, 0x190A87F0, 0xC47362BA, 0xB6B5242A]


def test_packed_draws_use_every_bit_of_six_calls_once(oracle):
    key, sweep, stream, tagc = 0x0123456789ABCDEF, 11, 3, oracle.TAG_SWEEP_PACKED | 1
    w = np.concatenate([oracle.philox4x32([9, sweep, stream, tagc | (call << 8)], [key & 0xFFFFFFFF, key >> 32]) for call in range(6)])
    bits = "".join(format(int(x), "032b") for x in w)
    got = [oracle.lib().orc_draw24_packed(key, 9 * 32 + b, sweep, stream, tagc) for b in range(32)]
    for g in range(8):
        a, b_, c = (int(w[3 * g + j]) for j in range(3))
        assert got[4 * g:4 * g + 4] == [a >> 8, b_ >> 8, c >> 8, ((a & 255) << 16) | ((b_ & 255) << 8) | (c & 255)]
    assert sum(bin(d).count("1") for d in got) == bits.count("1")  # a permutation of the 768 generated bits


def test_system_quad_mapping_one_call_serves_four_systems_at_a_site(oracle):
    """RNG-SPEC system-quad mapping (kernels that keep the same site of 32 systems in one word): counter = {colour rank, sweep,
    system >> 2, tag | colour}, system s draws out[s & 3] >> 8.  A Metropolis sweep of the oracle in that mode equals a replay of
    the log-form rule (sweep.rs:35-48, 247-257) in numpy f32 with draws taken straight from the generator."""
    shape, S, sweep = (4, 6), 7, 5
    lat = oracle.Lattice(shape)
    N, z = lat.n_spins, 2
    rng = np.random.default_rng(2)
    J = rng.standard_normal((N, z)).astype(np.float32)
    temps = np.linspace(0.7, 2.0, S).astype(np.float32)
    ids = rng.permutation(S).astype(np.int64)
    colour = (np.indices(shape).sum(axis=0) % 2).astype(np.uint16).reshape(-1)
    key = 0x0123456789ABCDEF
    spins = (2 * rng.integers(0, 2, size=(S, N)) - 1).astype(np.int8)
    got = spins.copy()
    lat.sweep_philox(got, J, temps, ids, colour, key, sweep, oracle.SWEEP_METROPOLIS, use_lookup=False, stream_is_slot=4)
    want = spins.copy()
    rank = np.zeros(N, np.int64)
    for c in range(2):
        rank[colour == c] = np.arange(int((colour == c).sum()))
    for slot in range(S):
        sys, s = int(ids[slot]), want[int(ids[slot])]
        for c in range(2):
            for i in np.flatnonzero(colour == c):
                h = np.float32(0.0)
                for d in range(z):  # sweep.rs:10-17: forward then backward, f32
                    h = np.float32(h + np.float32(s[lat.fwd(i, d)]) * J[i, d])
                    jb = lat.bwd(i, d)
                    h = np.float32(h + np.float32(s[jb]) * J[jb, d])
                out = oracle.philox4x32([int(rank[i]), sweep, sys >> 2, oracle.TAG_SWEEP_SYSQ | c], [key & 0xFFFFFFFF, key >> 32])
                u = np.float32(int(out[sys & 3]) >> 8) / np.float32(1 << 24)
                ec = np.float32(-np.float32(s[i]) * h)
                with np.errstate(divide="ignore"):
                    thr = np.float32(temps[slot] / np.float32(2.0)) * np.log(u, dtype=np.float32)
                if ec >= thr:
                    s[i] = -s[i]
    assert np.array_equal(got, want)
    assert not np.array_equal(got, spins)


def test_draw24_serves_four_indices_per_call(oracle):
    key = 0x0123456789ABCDEF
    out = oracle.philox4x32([5, 7, 9, oracle.TAG_SWEEP | 1], [key & 0xFFFFFFFF, key >> 32])
    for lane in range(4):
        assert oracle.lib().orc_draw24(key, 20 + lane, 7, 9, oracle.TAG_SWEEP | 1) == int(out[lane]) >> 8


# ---------------------------------------------------------------- lattice (geometry/lattice.rs:116-184)
def test_2d_neighbors(oracle):
    lat = oracle.Lattice([3, 4])
    assert lat.n_spins == 12 and lat.strides() == [4, 1]
    assert (lat.fwd(0, 0), lat.fwd(0, 1)) == (4, 1)
    assert (lat.bwd(0, 0), lat.bwd(0, 1)) == (8, 3)
    assert (lat.fwd(11, 0), lat.fwd(11, 1)) == (3, 8)


def test_3d_neighbors(oracle):
    lat = oracle.Lattice([2, 3, 4])
    assert lat.n_spins == 24 and lat.strides() == [12, 4, 1]
    assert [lat.fwd(0, d) for d in range(3)] == [12, 4, 1]


def test_triangular_neighbors(oracle):
    lat = oracle.Lattice([4, 4], [[1, 0], [0, 1], [1, -1]])
    assert lat.n_neighbors == 3 and lat.n_spins == 16
    assert [lat.fwd(0, d) for d in range(3)] == [4, 1, 7]
    assert [lat.bwd(0, d) for d in range(3)] == [12, 3, 13]
    assert lat.fwd(5, 2) == 8 and lat.bwd(5, 2) == 2
    assert [lat.fwd(15, d) for d in range(3)] == [3, 12, 2]


def test_l2_forward_equals_backward_and_both_are_counted(oracle):
    # L=2 -> fwd == bwd neighbour, two distinct bonds (sweep.rs:11-17)
    lat = oracle.Lattice([2, 2])
    for i in range(4):
        for d in range(2):
            assert lat.fwd(i, d) == lat.bwd(i, d)


# ---------------------------------------------------------------- energy (spins/energy.rs:117-147)
def test_energy_and_magnetisation_kat(oracle):
    lat = oracle.Lattice([2, 3])
    J = np.ones((6, 2), dtype=np.float32)
    spins = np.array([[1, 1, 1, 1, 1, 1], [1, -1, 1, -1, 1, -1]], dtype=np.int8)
    e, m = lat.energies_mags(spins, J)
    assert m.tolist() == [6, 0]
    fwd = lat.fwd_table()
    for r in range(2):
        inter = [float(spins[r, i]) * float(spins[r, fwd[i, d]]) * 1.0 for i in range(6) for d in range(2)]
        assert e[r] == np.float32(np.float32(sum(inter)) / np.float32(6))
    assert e[0] == np.float32(2.0)


# ---------------------------------------------------------------- LUT (mcmc/sweep.rs:346-380)
def test_unit_lookup_is_fail_closed(oracle):
    ok = oracle.metropolis_lookup
    assert ok([-1.0, 0.0, 1.0], [0.5, 2.0], 2) is not None
    assert ok([-2.0, 2.0], [1.0], 2) is None
    assert ok([-1.0, np.nan], [1.0], 2) is None
    assert ok([-1.0, 1.0], [0.0], 2) is None
    assert ok([-1.0, 1.0], [np.frombuffer(np.uint32(1).tobytes(), dtype=np.float32)[0]], 2) is None


def test_unit_lookup_cutoffs_match_legacy_boundaries(oracle):
    temps = np.array([0.7, 2.0, 5.0], dtype=np.float32)
    table = oracle.metropolis_lookup([-1.0, 0.0, 1.0], temps, 3)
    L = oracle.lib()
    for t, temp in enumerate(temps):
        for ec in range(-6, 7):
            count = int(table[t, ec + 6])
            if count > 0:
                assert L.orc_metropolis_legacy_accepts(temp, ec, count - 1)
            if count < F24:
                assert not L.orc_metropolis_legacy_accepts(temp, ec, count)
            if ec >= 0:
                assert count == F24


def test_lookup_example_row_matches_exp_rule(oracle):
    # SURVEY §3.2 example: P(accept) = exp(2 ec / T)
    table = oracle.metropolis_lookup([1.0], [0.8], 3)
    for ec in range(-6, 1):
        assert abs(int(table[0, ec + 6]) - F24 * np.exp(2 * ec / np.float32(0.8))) <= 2


@pytest.mark.parametrize("temp", [0.5, 1.0, 2.269, 7.5])
def test_gibbs_counts_match_log_rule_boundaries(oracle, temp):
    L = oracle.lib()
    for ec in range(-6, 7):
        count = L.orc_gibbs_accepted_count(temp, ec)
        if count > 0:
            assert L.orc_gibbs_legacy_accepts(temp, ec, count - 1)
        if count < F24:
            assert not L.orc_gibbs_legacy_accepts(temp, ec, count)
        # heat-bath probability 1/(1+exp(-2 ec/T))
        assert abs(count / F24 - 1.0 / (1.0 + np.exp(-2.0 * ec / temp))) < 1e-6


def test_gibbs_rule_is_monotone_on_a_dense_draw_sample(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for temp, ec in [(1.1, -2), (3.64, 4), (0.9, 0)]:
        count = L.orc_gibbs_accepted_count(temp, ec)
        draws = np.concatenate([rng.integers(0, F24, 4000), np.arange(max(count - 300, 0), min(count + 300, F24))])
        for d in draws:
            assert bool(L.orc_gibbs_legacy_accepts(temp, ec, int(d))) == (d < count)


def _lookup_vs_log(oracle, lat, mode):
    n = lat.n_spins
    z = lat.n_neighbors
    J = np.array([(-1.0, 0.0, 1.0)[i % 3] for i in range(n * z)], dtype=np.float32)
    temps = np.array([0.7, 2.0, 5.0], dtype=np.float32)
    sid = [2, 0, 1]
    init = np.array([-1 if i % 5 == 0 else 1 for i in range(3 * n)], dtype=np.int8).reshape(3, n)
    if mode == "xoshiro":
        def rngs():
            st = np.zeros((3, 4), dtype=np.uint64)
            for k, seed in enumerate((71, 72, 73)):
                oracle.lib().orc_xoshiro_seed_from_u64(st[k].ctypes.data, seed)
            return st
        a, b, ra, rb = init.copy(), init.copy(), rngs(), rngs()
        for _ in range(20):
            lat.sweep_xoshiro(a, J, temps, sid, ra, oracle.SWEEP_METROPOLIS, True)
            lat.sweep_xoshiro(b, J, temps, sid, rb, oracle.SWEEP_METROPOLIS, False)
        assert np.array_equal(ra, rb)
    else:
        shape = lat.shape
        colour = (np.indices(tuple(shape)).sum(axis=0) % 2).astype(np.uint16).reshape(-1)
        assert lat.colouring_is_valid(colour)
        a, b = init.copy(), init.copy()
        for sweep in range(20):
            for sm in (oracle.SWEEP_METROPOLIS, oracle.SWEEP_GIBBS):
                lat.sweep_philox(a, J, temps, sid, colour, 99, 2 * sweep + sm, sm, use_lookup=True)
                lat.sweep_philox(b, J, temps, sid, colour, 99, 2 * sweep + sm, sm, use_lookup=False)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, init)


def test_unit_lookup_matches_log_with_permuted_systems(oracle):
    # mcmc/sweep.rs:382-442
    _lookup_vs_log(oracle, oracle.Lattice([8, 8]), "xoshiro")
    _lookup_vs_log(oracle, oracle.Lattice([8, 8], [[1, 0], [0, 1]]), "xoshiro")


def test_philox_lookup_matches_log_form_metropolis_and_gibbs(oracle):
    _lookup_vs_log(oracle, oracle.Lattice([8, 8]), "philox")
    _lookup_vs_log(oracle, oracle.Lattice([4, 6, 4]), "philox")


# ---------------------------------------------------------------- PT (mcmc/tempering.rs:110-138)
def test_full_ladder_attempts_every_edge_in_requested_parity_order(oracle):
    assert oracle.full_ladder_edges(5, 0) == [0, 2, 1, 3]
    assert oracle.full_ladder_edges(5, 1) == [1, 3, 0, 2]
    assert oracle.full_ladder_edges(1, 0) == []
    assert oracle.full_ladder_edges(2, 1) == [0]


def test_tracks_hot_cold_hot_across_attempts(oracle):
    # simulation/realization.rs:285-302
    ea, eacc, rt = oracle.pt_replay(1, [0.5, 1.0, 2.0], [(1, True, 1, 2), (0, True, 0, 2), (0, True, 2, 0), (1, True, 2, 1)])
    assert ea.tolist() == [2, 2] and eacc.tolist() == [2, 2]
    assert rt[2] == 1 and rt.sum() == 1


# ---------------------------------------------------------------- realization (realization.rs:267-283) + interface
@pytest.mark.parametrize("mode", ["xoshiro", "philox"])
def test_reset_replays_spins_and_pt_state(oracle, mode):
    J = np.ones((3, 3, 2), dtype=np.float32)
    kw = {}
    if mode == "philox":
        # 3x3 has odd extents: a proper colouring needs 3 colours, c = (x + y) mod 3
        kw = dict(rng_mode=oracle.RNG_PHILOX, colour=(np.indices((3, 3)).sum(axis=0) % 3))
    sim = oracle.Sim([3, 3], J, [1.0, 2.0], n_replicas=2, seed=17, **kw)
    first = sim.spins()
    assert set(np.unique(first)) <= {-1, 1}
    r = sim.sample(3, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    assert np.all(r["per_disorder"]["parallel_tempering"]["edge_attempts"] == 6)
    sim.reset()
    assert np.array_equal(sim.spins(), first)
    assert sim.system_ids().tolist() == [0, 1, 2, 3]
    r = sim.sample(1, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    assert np.all(r["per_disorder"]["parallel_tempering"]["edge_attempts"] == 2)
    sim.reset(seed=99)
    other = sim.spins()
    sim.reset(seed=99)
    assert np.array_equal(sim.spins(), other) and not np.array_equal(other, first)


def test_pt_counters_accumulate_across_calls(oracle):
    # tests/test_sampling_interfaces.py:75-118 (PT part)
    rng = np.random.default_rng(11)
    J = (2 * rng.integers(0, 2, size=(4, 4, 2)) - 1).astype(np.float32)
    sim = oracle.Sim([4, 4], J, [1.0, 2.0, 4.0], n_replicas=2, seed=11)
    r = sim.sample(2, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    pt = r["per_disorder"]["parallel_tempering"]
    assert pt["edge_attempts"].shape == (1, 2) and np.all(pt["edge_attempts"] == 4)
    assert pt["round_trips"].shape == (1, 2, 3)
    r = sim.sample(1, pt_interval=1, pt_schedule="full_ladder", warmup_ratio=0)
    assert np.all(r["per_disorder"]["parallel_tempering"]["edge_attempts"] == 6)


def test_result_keys_shapes_and_disorder_aggregation(oracle):
    rng = np.random.default_rng(3)
    J = (2 * rng.integers(0, 2, size=(3, 4, 4, 2)) - 1).astype(np.float32)
    temps = np.array([1.0, 2.0], dtype=np.float32)
    sim = oracle.Sim([4, 4], J, temps, n_replicas=2, seed=5)
    r = sim.sample(8, pt_interval=2, warmup_ratio=0.25)
    for k in ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4",
              "link_overlap", "link_overlap2", "link_overlap4"):
        assert r[k].shape == (2,) and r[k].dtype == np.float64
    assert r["overlap_histogram"].shape == (2, 17)
    assert r["per_sample_overlap_histogram"].shape == (3, 2, 17)
    # 6 recorded sweeps x 1 pair per sample
    assert np.all(r["per_sample_overlap_histogram"].sum(axis=2) == 6)
    assert np.array_equal(r["per_sample_overlap_histogram"].sum(axis=0), r["overlap_histogram"])
    # disorder mean == mean of single-realization runs (results.rs:165-180, 250-259)
    singles = [oracle.Sim([4, 4], J[d], temps, n_replicas=2, seed=5) for d in range(3)]
    # realization seeds differ by index (lib.rs:158-164): rebuild sample d as realization 0 would differ,
    # so only check the exact arithmetic identity on energies2 >= energies^2 and mags2 bounds here
    assert np.all(r["energies2"] >= r["energies"] ** 2 - 1e-12)
    assert np.all((r["mags2"] >= 0) & (r["mags2"] <= 1))
    del singles


def test_threads_do_not_change_results(oracle):
    rng = np.random.default_rng(8)
    J = rng.standard_normal((6, 4, 4, 4, 3)).astype(np.float32)
    temps = np.linspace(0.8, 1.8, 4).astype(np.float32)
    a = oracle.Sim([4, 4, 4], J, temps, n_replicas=2, seed=1).sample(10, pt_interval=1, n_threads=1)
    b = oracle.Sim([4, 4, 4], J, temps, n_replicas=2, seed=1).sample(10, pt_interval=1, n_threads=4)
    for k in a:
        if k != "per_disorder":
            assert np.array_equal(a[k], b[k]), k


def test_invalid_config_is_rejected(oracle):
    sim = oracle.Sim([4, 4], np.ones((4, 4, 2), np.float32), [2.0])
    before = sim.spins()
    with pytest.raises(ValueError, match="n_sweeps must be >= 1"):
        sim.sample(0)
    assert np.array_equal(sim.spins(), before)


def test_ferromagnet_orders_below_tc_and_not_above(oracle):
    # physics sanity for the reference-faithful mode: 2-D Ising, T_c = 2.269
    J = np.ones((16, 16, 2), dtype=np.float32)
    r = oracle.Sim([16, 16], J, [1.5, 3.5], seed=42).sample(600, warmup_ratio=0.5)
    assert r["mags2"][0] > 0.9 and r["mags2"][1] < 0.1
    # exact 2-D energy at T=1.5: e = +sum J s s / N ~ 1.951 (Onsager), sign convention energy.rs:103-108
    assert abs(r["energies"][0] - 1.951) < 0.01


# ---- statistics/autocorrelation.rs:214-374 (the reference's own unit tests of the accumulator) ----
def _deterministic_values(sample):  # autocorrelation.rs:283-288
    return [float(np.float32((sample * 13 % 31)) / np.float32(8.0) - np.float32(2.0)),
            float(np.float32((sample * 7 % 23)) / np.float32(4.0) - np.float32(1.5))]


def _brute_force_gamma(series, max_lag):  # autocorrelation.rs:290-308
    series = np.asarray(series, dtype=np.float64)
    count = len(series)
    mean = series.sum() / count
    variance = (series * series).sum() / count - mean * mean
    out = []
    for delta in range(max_lag + 1):
        pairs = max(count - delta, 0)
        if pairs == 0 or variance <= 0.0:
            out.append(1.0 if delta == 0 else 0.0)
            continue
        product_sum = float(np.sum(series[delta:] * series[:count - delta]))
        out.append((product_sum / pairs - mean * mean) / variance)
    return np.asarray(out)


def test_autocorr_empty_and_constant_series_are_degenerate(oracle):  # autocorrelation.rs:331-343
    assert np.array_equal(oracle.autocorr_gamma(np.zeros((0, 1)), 4), [[1.0, 0.0, 0.0, 0.0, 0.0]])
    assert np.array_equal(oracle.autocorr_gamma(np.full((8, 1), 3.5), 4), [[1.0, 0.0, 0.0, 0.0, 0.0]])


def test_autocorr_ring_matches_brute_force_across_wraps(oracle):  # autocorrelation.rs:310-329, 345-373
    for n, lag in ((41, 7), (128, 40)):
        values = np.asarray([_deterministic_values(s) for s in range(n)])
        gamma = oracle.autocorr_gamma(values, lag)
        for t in range(2):
            np.testing.assert_allclose(gamma[t], _brute_force_gamma(values[:, t], lag), rtol=0, atol=1e-10)
        assert gamma[0][0] == pytest.approx(1.0, abs=1e-12)


def test_sokal_tau_window(oracle):  # autocorrelation.rs:201-210
    assert oracle.sokal_tau([1.0]) == 0.5
    assert oracle.sokal_tau([1.0, 0.0, 0.0]) == 0.5                    # w = 1 >= 5 * 0.5 fails, w = 2 ... runs out
    g = [1.0] + [0.5 ** w for w in range(1, 40)]
    tau = oracle.sokal_tau(g)
    assert 1.4 < tau < 1.5                                              # 0.5 + sum of the geometric tail up to the window
    assert oracle.sokal_tau([1.0, 0.05, 0.9, 0.9]) == pytest.approx(2.35)        # window never closes: the whole sum
    assert oracle.sokal_tau([1.0, 0.1, 0.05, 0.0, 0.0, 5.0]) == pytest.approx(0.65)  # closes at w = 4 >= 5 * 0.65: the tail is ignored


def test_autocorrelation_keys_and_lag_clamp(oracle):
    """mod.rs:342-344: the lag is clamped to recorded / 4 (at least 1); taus are means over realizations (results.rs:217-272)."""
    rng = np.random.default_rng(0)
    J = (2 * rng.integers(0, 2, size=(2, 4, 4, 2)) - 1).astype(np.float32)
    sim = oracle.Sim((4, 4), J, np.asarray([1.5, 2.5], np.float32), n_replicas=2, seed=3)
    res = sim.sample(40, "metropolis", autocorrelation_max_lag=1000)
    assert res["mags2_tau"].shape == (2,) and res["overlap2_tau"].shape == (2,)
    assert np.all(np.isfinite(res["mags2_tau"])) and np.all(np.isfinite(res["overlap2_tau"]))
    assert np.allclose(res["mags2_tau"], sim.last_per_sample_taus[:, 0, :].sum(axis=0) / 2)
    sim1 = oracle.Sim((4, 4), J[0], np.asarray([1.5, 2.5], np.float32), n_replicas=1, seed=3)
    assert "overlap2_tau" not in sim1.sample(40, "metropolis", autocorrelation_max_lag=5)


# ---- clusters/fk.rs:28-171 restated under RNG-SPEC draws (oracle.fk_update) ----
def _fk(oracle, shape, spins, J, temperature, wolff, key=0x1234ABCD5678EF01, sweep_index=7, system_id=3):
    import ctypes as C

    lat = oracle.Lattice(shape)
    s = np.ascontiguousarray(spins, dtype=np.int8).copy()
    Jc = np.ascontiguousarray(J, dtype=np.float32)
    f = oracle.lib().orc_fk_update
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
    f(lat.h, s.ctypes.data, Jc.ctypes.data, float(temperature), key, sweep_index, system_id, int(wolff))
    return s.reshape(-1)


def test_fk_bond_count_limits(oracle):
    f = oracle.lib().orc_fk_bond_count
    import ctypes as C
    f.restype, f.argtypes = C.c_uint32, [C.c_float]
    assert f(1e-3) == 1 << 24                                  # T -> 0: every satisfied bond is activated
    assert f(1e9) == 0 or f(1e9) < 4                           # T -> inf: none
    p = 1.0 - np.exp(np.float32(-2.0) / np.float32(2.269))
    assert abs(f(2.269) / 2.0**24 - float(p)) < 2.0**-23       # fk.rs:113


def test_fk_cold_limit_flips_whole_domains(oracle):
    """T -> 0: clusters = connected domains of satisfied bonds.  A ferromagnet with one flipped 2x2 block has two domains;
    Swendsen-Wang flips each as a whole (or not), Wolff flips exactly the seed's domain."""
    shape = (6, 6)
    spins = np.ones(shape, np.int8)
    spins[1:3, 1:3] = -1
    J = np.ones(shape + (2,), np.float32)
    inside = spins.reshape(-1) == -1
    for wolff in (0, 1):
        seen = set()
        for sweep in range(48 if wolff else 12):  # the Wolff seed lands in the small domain with probability 1/9
            out = _fk(oracle, shape, spins, J, 1e-3, wolff, sweep_index=sweep)
            a, b = out[inside] * spins.reshape(-1)[inside], out[~inside] * spins.reshape(-1)[~inside]
            assert len(set(a)) == 1 and len(set(b)) == 1       # each domain moves as one
            if wolff:
                assert (a[0] == -1) != (b[0] == -1)            # exactly one domain flips
            seen.add((int(a[0]), int(b[0])))
        assert len(seen) >= 2                                  # the draws do differ from sweep to sweep


def test_fk_hot_limit_is_independent_coin_flips_and_antiferro_bonds_never_join(oracle):
    shape = (8, 8)
    spins = np.ones(shape, np.int8)
    J = np.ones(shape + (2,), np.float32)
    out = _fk(oracle, shape, spins, J, 1e9, 0)
    frac = np.mean(out == -1)
    assert 0.3 < frac < 0.7 and len(set(out.tolist())) == 2    # single-site clusters, fair coins (fk.rs:159-170)
    # all bonds unsatisfied (J = -1 on an aligned state): nothing joins even at T -> 0
    out2 = _fk(oracle, shape, spins, -J, 1e-3, 0)
    assert 0.3 < np.mean(out2 == -1) < 0.7
