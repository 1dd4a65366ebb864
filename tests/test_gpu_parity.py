"""T2/T3: the CUDA path (through the C ABI) against the oracle under the shared Philox draws (RNG-SPEC).
Bit-exact: spins, system_ids, +-J energies, magnetisations, overlap dots, histograms, PT counters, f64 means.
fp32 couplings: bit-exact spins with exact_log (host-libm log table), energies within 1e-5 relative."""
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.gpu

TRI = [[1, 0], [0, 1], [1, -1]]
FCC = [[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1]]


def couplings(kind, shape, z, D, seed):
    rng = np.random.default_rng(seed)
    full = (D,) + tuple(shape) + (z,)
    if kind == "ferro":
        J = np.ones(full, np.float32)
    elif kind == "bimodal":
        J = (2 * rng.integers(0, 2, size=full) - 1).astype(np.float32)
    elif kind == "diluted":  # {-1, 0, +1}: lookup-eligible with odd fields (sweep.rs:110-112)
        J = rng.integers(-1, 2, size=full).astype(np.float32)
    else:
        J = rng.standard_normal(full).astype(np.float32)
    return J[0] if D == 1 else J


def make_pair(oracle, shape, kind, temps, R, D, offsets=None, layout="int8", seed=1234, jseed=7):
    import peapods_b200 as pb

    z = len(shape) if offsets is None else len(offsets)
    J = couplings(kind, shape, z, D, jseed)
    temps = np.asarray(temps, np.float32)
    colour, _ = pb.colouring(shape, offsets)
    gpu = pb.IsingSimulation(list(shape), J, temps, R, offsets, seed, layout=layout)
    assert gpu.layout == layout
    # ferromagnets whose rows split into 64-site word pairs are kept as one bit per spin and draw through the packed mapping
    # fp32 couplings with >= 16 systems on rows of 32 k sites keep the same site of 32 systems in one word (system-quad mapping)
    mode = (oracle.RNG_PHILOX_MSC if layout == "msc" else oracle.RNG_PHILOX_PACKED if gpu.rows_packed
            else oracle.RNG_PHILOX_SYSQ if gpu.sys_words else oracle.RNG_PHILOX)
    cpu = oracle.Sim(shape, J, temps, n_replicas=R, offsets=offsets, seed=seed, rng_mode=mode, colour=colour)
    return gpu, cpu


def assert_state_equal(gpu, cpu, D):
    for d in range(D):
        assert np.array_equal(gpu.get_spins(d), cpu.spins(d)), f"spins differ, realization {d}"
        assert np.array_equal(gpu.get_system_ids(d), cpu.system_ids(d)), f"system_ids differ, realization {d}"


def assert_results_equal(rg, rc, exact=True, rtol=0.0):
    keys = [k for k in rc if k != "per_disorder"]
    for k in keys:
        a, b = rg[k], rc[k]
        if k == "overlap_histogram":
            a = np.stack(a)
        if exact:
            assert np.array_equal(np.asarray(a), np.asarray(b)), k
        else:
            np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=1e-7, err_msg=k)
    if "per_disorder" in rc:
        for k, v in rc["per_disorder"]["parallel_tempering"].items():
            assert np.array_equal(rg["per_disorder"]["parallel_tempering"][k], v), k


CASES_INT8 = [
    # shape, kind, offsets, temps, R, D
    ((8, 8), "ferro", None, [1.5, 2.27, 3.0], 2, 1),
    ((32, 32), "ferro", None, np.linspace(1.5, 3.0, 16), 2, 1),        # BASELINE config 1 geometry
    ((8, 8), "bimodal", None, [0.7, 2.0, 5.0], 2, 3),
    ((4, 4, 4), "bimodal", None, [0.8, 1.1, 1.4], 4, 2),
    ((8, 8, 8), "bimodal", None, np.linspace(0.8, 1.4, 4), 2, 2),
    ((16, 16, 16), "bimodal", None, [0.8, 1.4], 2, 1),                  # BASELINE config 2 geometry
    ((6, 6), "ferro", TRI, [3.0, 3.64, 4.2], 2, 1),
    ((8, 8), "bimodal", TRI, [1.0, 3.64], 2, 2),
    ((2, 2), "bimodal", None, [1.0, 2.0], 2, 2),                        # L=2: fwd == bwd neighbour
    ((2, 2, 2), "ferro", None, [4.5], 2, 1),
    ((5, 7), "bimodal", None, [0.8, 2.5], 2, 1),                        # odd extents: greedy colouring
    ((3, 3), "diluted", None, [1.0, 2.0], 2, 2),
    ((6, 6), "diluted", None, [0.9, 1.7, 2.9], 3, 2),                   # zeros -> odd local fields
    ((4, 4, 4), "bimodal", FCC, [2.0, 9.8], 2, 1),
    ((4, 6, 2, 2), "bimodal", None, [1.5, 6.0], 2, 1),                  # 4-D
]


@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", CASES_INT8)
@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
def test_int8_trajectory_is_bit_exact(oracle, shape, kind, offsets, temps, R, D, mode):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    assert_state_equal(gpu, cpu, D)  # INIT domain
    for n_sweeps in (1, 2, 17):
        rg = gpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        rc = cpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)


@pytest.mark.parametrize("schedule", ["single_random_edge", "full_ladder"])
@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", [
    ((8, 8), "ferro", None, np.linspace(1.5, 3.0, 6), 2, 1),
    ((4, 4, 4), "bimodal", None, np.linspace(0.8, 1.6, 5), 4, 3),
    ((6, 6), "bimodal", TRI, [1.0, 1.5, 2.5], 3, 2),
    ((4, 4), "bimodal", None, [1.3], 2, 2),                            # single temperature: PT is a no-op
])
def test_int8_sample_with_pt_and_overlap_is_bit_exact(oracle, shape, kind, offsets, temps, R, D, schedule):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    for n_sweeps, interval in ((100, 1), (23, 3)):
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        for d in range(D):
            assert np.array_equal(gpu.get_energies(d), cpu.energies(d))


@pytest.mark.parametrize("shape,D,R,temps", [
    ((4, 4, 4), 32, 2, [0.8, 1.1, 1.4]),
    ((4, 4, 4), 45, 4, [0.9, 1.3]),            # padded second word group
    ((8, 8, 8), 64, 2, np.linspace(0.8, 1.4, 4)),
    ((8, 8), 40, 2, [0.7, 2.0, 5.0]),          # 2-D: 4 bond words per site
    ((16, 16, 16), 32, 2, [0.8, 1.4]),
])
@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
def test_msc_trajectory_is_bit_exact(oracle, shape, D, R, temps, mode):
    gpu, cpu = make_pair(oracle, shape, "bimodal", temps, R, D, layout="msc")
    assert_state_equal(gpu, cpu, D)
    for n_sweeps in (1, 2, 9):
        rg = gpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        rc = cpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)


@pytest.mark.parametrize("schedule", ["single_random_edge", "full_ladder"])
@pytest.mark.parametrize("shape,kind,offsets,D,R,temps", [
    ((4, 4, 4), "bimodal", None, 40, 4, np.linspace(0.8, 1.6, 5)),
    ((8, 8, 8), "bimodal", None, 32, 2, np.linspace(0.8, 1.4, 6)),
    ((6, 6), "bimodal", TRI, 33, 2, [1.0, 1.5, 2.5]),
    ((8, 8), "ferro", None, 32, 2, [1.5, 2.27, 3.0]),
])
def test_msc_sample_with_pt_and_overlap_is_bit_exact(oracle, shape, kind, offsets, D, R, temps, schedule):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets, layout="msc")
    for n_sweeps, interval in ((60, 1), (23, 3)):
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)


MSC3D_CASES = [
    # shape, kind, D, R, temps  -- all taken by the stride-based 3-D multispin kernel
    ((8, 8, 8), "bimodal", 32, 1, [0.8, 1.1, 1.4]),
    ((8, 8, 8), "bimodal", 45, 4, np.linspace(0.8, 1.6, 5)),          # padded second word group
    ((4, 6, 8), "bimodal", 33, 2, [0.9, 1.3, 2.0]),                   # non-cubic, one quad per row
    ((2, 2, 8), "bimodal", 32, 4, [1.0, 2.0]),                        # extent 2: forward == backward neighbour
    ((6, 4, 16), "bimodal", 64, 2, [0.8, 1.2]),
    ((8, 8, 8), "ferro", 32, 2, [3.5, 4.5, 5.5]),                     # no coupling words
    ((16, 16, 16), "bimodal", 64, 4, np.linspace(0.8, 1.4, 3)),       # BASELINE config 2 geometry, R = 4 (in-sweep energy path)
    ((16, 16, 16), "bimodal", 33, 2, [0.8, 1.1, 1.4, 1.7]),           # the same path with one replica pair, padded group
    ((16, 16, 16), "ferro", 32, 4, [4.2, 4.5]),
    ((16, 16, 16), "bimodal", 32, 1, [0.9, 1.3]),                     # R = 1: epilogue walks the spins for E and M
]


@pytest.mark.parametrize("schedule", ["single_random_edge", "full_ladder"])
@pytest.mark.parametrize("shape,kind,D,R,temps", MSC3D_CASES)
def test_msc3d_sample_is_bit_exact(oracle, shape, kind, D, R, temps, schedule):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, layout="msc")
    assert gpu.uses_msc3d
    assert_state_equal(gpu, cpu, D)
    for n_sweeps, interval, mode in ((1, None, "metropolis"), (24, 1, "metropolis"), (11, 3, "metropolis"), (7, 2, "gibbs")):
        rg = gpu.sample(n_sweeps, mode, pt_interval=interval, pt_schedule=schedule)
        rc = cpu.sample(n_sweeps, mode, pt_interval=interval, pt_schedule=schedule)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        for d in range(0, D, 7):
            assert np.array_equal(gpu.get_energies(d), cpu.energies(d))


@pytest.mark.parametrize("streams,groups,macro", [(2, 1, 4), (3, 2, 16), (1, 1, 1)])
def test_msc3d_chunked_multistream_execution_is_bit_exact(oracle, monkeypatch, streams, groups, macro):
    """The chunk / stream / macro-batch plan of pp_sample must not change any result."""
    monkeypatch.setenv("PP_STREAMS", str(streams))
    monkeypatch.setenv("PP_CHUNK_GROUPS", str(groups))
    monkeypatch.setenv("PP_MACRO_BATCH", str(macro))
    shape, temps, R, D = (4, 4, 8), np.linspace(0.8, 1.6, 5), 4, 150   # 5 word groups, the last one padded
    gpu, cpu = make_pair(oracle, shape, "bimodal", temps, R, D, layout="msc")
    assert gpu.uses_msc3d
    seen = []
    for n_sweeps, interval, schedule in ((37, 1, "single_random_edge"), (21, 2, "full_ladder"), (9, None, "full_ladder")):
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule, on_sweep=seen.append)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
    assert seen == list(range(37)) + list(range(21)) + list(range(9))


def test_shards_with_sample_offset_equal_the_unsharded_run(oracle):
    """Multi-GPU sharding = handles over blocks of realizations with their global sample_offset (one per GPU in
    production; here both on cuda:0).  Merged through peapods_b200.sharded it must equal the single handle."""
    import peapods_b200 as pb
    from peapods_b200.sharded import merge_results, shard_bounds

    shape, temps, R, D = (4, 4, 8), np.linspace(0.8, 1.6, 4).astype(np.float32), 2, 96
    J = couplings("bimodal", shape, 3, D, 3)
    whole = pb.IsingSimulation(list(shape), J, temps, R, None, 99, layout="msc")
    kw = dict(pt_interval=1, pt_schedule="single_random_edge")
    ref = whole.sample(25, "metropolis", **kw)
    parts = []
    for rank in range(2):
        first, count = shard_bounds(D, 2, rank)
        sim = pb.IsingSimulation(list(shape), J[first:first + count], temps, R, None, 99, layout="msc", sample_offset=first)
        res = sim.sample(25, "metropolis", **kw)
        parts.append({"result": res, "per_sample_means": sim.last_per_sample_means, "n_replicas": R})
        for d in range(count):
            assert np.array_equal(sim.get_spins(d), whole.get_spins(first + d))
    merged = merge_results(parts)
    for k in ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4", "link_overlap2"):
        assert np.array_equal(merged[k], ref[k]), k
    assert np.array_equal(np.stack(merged["overlap_histogram"]), np.stack(ref["overlap_histogram"]))
    assert np.array_equal(merged["per_sample_overlap_histogram"], ref["per_sample_overlap_histogram"])
    np.testing.assert_allclose(merged["ql_at_q_sum"], ref["ql_at_q_sum"], rtol=1e-12)
    for k, v in ref["per_disorder"]["parallel_tempering"].items():
        assert np.array_equal(merged["per_disorder"]["parallel_tempering"][k], v), k


def test_msc3d_operator_entry_points_match_oracle(oracle):
    shape, temps, R, DD = (4, 4, 8), [0.9, 1.2, 1.5], 4, 37
    gpu, cpu = make_pair(oracle, shape, "bimodal", temps, R, DD, layout="msc")
    assert gpu.uses_msc3d
    lat = oracle.Lattice(shape)
    J = couplings("bimodal", shape, 3, DD, 7)
    rng = np.random.default_rng(11)
    S, N, T = R * len(temps), lat.n_spins, len(temps)
    spins = (2 * rng.integers(0, 2, size=(DD, S, N)) - 1).astype(np.int8)
    for d in range(DD):
        gpu.set_spins(spins[d], d)
    e_g, m_g = gpu.op_energies_mags()
    ds, dl = gpu.op_overlap()
    for d in range(DD):
        e_c, m_c = lat.energies_mags(spins[d], J[d])
        assert np.array_equal(m_g[d], m_c) and np.array_equal(e_g[d], e_c)
        for p in range(R // 2):
            for t in range(T):
                assert (ds[d, p, t], dl[d, p, t]) == lat.overlap_dots(spins[d][(2 * p) * T + t], spins[d][(2 * p + 1) * T + t])


@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
@pytest.mark.parametrize("shape,offsets,D", [((4, 4, 4), None, 2), ((8, 8, 8), None, 1), ((6, 6), TRI, 2)])
def test_gaussian_couplings_exact_log_spins_and_energy_tolerance(oracle, shape, offsets, D, mode):
    temps = [0.8, 1.3, 1.8]
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, 2, D, offsets)
    for n_sweeps in (1, 12):
        rg = gpu.sample(n_sweeps, mode, warmup_ratio=0.25, exact_log=True)
        rc = cpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        assert_state_equal(gpu, cpu, D)  # same spins => the energy comparison is on identical states
        e_g, m_g = gpu.op_energies_mags()
        for d in range(D):
            lat = oracle.Lattice(shape, offsets)
            Jd = gpu_couplings(shape, offsets, D)[d]
            e_c, m_c = lat.energies_mags(cpu.spins(d).reshape(-1, lat.n_spins), Jd)
            np.testing.assert_allclose(e_g[d], e_c, rtol=1e-5, atol=1e-6)
            assert np.array_equal(m_g[d], m_c)
        # integer observables are exact, f32-energy-derived ones within tolerance
        for k in ("mags", "mags2", "mags4", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
            assert np.array_equal(rg[k], rc[k]), k
        np.testing.assert_allclose(rg["energies"], rc["energies"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(rg["energies2"], rc["energies2"], rtol=2e-5, atol=1e-7)


SYS_WORD_CASES = [
    # shape, offsets, temps, R, D: fp32 couplings, two colours, rows of 32 k sites, >= 16 systems per realization
    ((4, 4, 32), None, np.linspace(0.8, 1.8, 4), 4, 2),       # 16 systems: half a word
    ((2, 6, 32), None, np.linspace(0.8, 1.8, 5), 4, 1),       # 20 systems, L0 = 2 (forward == backward neighbour row)
    ((4, 32), None, np.linspace(0.7, 2.4, 9), 4, 3),          # 2-D, 36 systems: one full word + four lanes of a second
    ((4, 2, 64), None, np.linspace(0.8, 1.8, 8), 4, 1),       # rows of two words, exactly one word of systems
    ((6, 4, 32), None, np.linspace(0.8, 1.8, 11), 6, 2),      # 66 systems: three word slabs, three replica pairs
    ((8, 8, 8), None, np.linspace(0.8, 1.8, 8), 2, 2),        # rows of 8 sites: four rows per word of the transposed view
    ((4, 6, 16), None, np.linspace(0.8, 1.8, 6), 4, 1),       # rows of 16 sites
    ((4, 4, 24), None, np.linspace(0.8, 1.8, 5), 4, 1),       # rows of 24 sites straddle the words of the transposed view
    ((12, 16), None, np.linspace(0.7, 2.4, 8), 2, 2),         # 2-D, rows of 16 sites
]


@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
@pytest.mark.parametrize("shape,offsets,temps,R,D", SYS_WORD_CASES)
def test_sys_word_trajectory_exact_log_spins_and_energy_tolerance(oracle, shape, offsets, temps, R, D, mode):
    """fp32 couplings with the same site of 32 systems in one word (pp_kernels_swords.cuh) against the oracle in the system-quad
    draw mapping: spins, magnetisation and overlap observables bit for bit, f32 energies within 1e-5 relative."""
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, R, D, offsets)
    assert gpu.sys_words
    assert_state_equal(gpu, cpu, D)  # INIT domain through the pack / unpack of the int8 view
    lat = oracle.Lattice(shape, offsets)
    J = couplings("gaussian", shape, len(shape), D, 7).reshape((D, -1))
    for n_sweeps in (1, 2, 13):
        rg = gpu.sample(n_sweeps, mode, warmup_ratio=0.25, exact_log=True)
        rc = cpu.sample(n_sweeps, mode, warmup_ratio=0.25)
        assert_state_equal(gpu, cpu, D)
        e_g, m_g = gpu.op_energies_mags()
        ds, dl = gpu.op_overlap()
        T = len(temps)
        for d in range(D):
            sp = cpu.spins(d).reshape(-1, lat.n_spins)
            e_c, m_c = lat.energies_mags(sp, J[d])
            np.testing.assert_allclose(e_g[d], e_c, rtol=1e-5, atol=1e-6)
            assert np.array_equal(m_g[d], m_c)
            ids = cpu.system_ids(d)
            for p in range(R // 2):
                for t in range(T):
                    assert (ds[d, p, t], dl[d, p, t]) == lat.overlap_dots(sp[ids[(2 * p) * T + t]], sp[ids[(2 * p + 1) * T + t]])
        for k in ("mags", "mags2", "mags4", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
            assert np.array_equal(rg[k], rc[k]), k
        assert np.array_equal(np.stack(rg["overlap_histogram"]), np.asarray(rc["overlap_histogram"]))
        np.testing.assert_allclose(rg["energies"], rc["energies"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(rg["energies2"], rc["energies2"], rtol=2e-5, atol=1e-7)


def test_sys_word_set_spins_round_trip_and_cluster_moves(oracle):
    """the int8 view of a system-word handle: set_spins / get_spins round trip, energies of a given configuration, and a
    Houdayer move through the view equal to the oracle's."""
    shape, temps, R, D = (4, 4, 32), np.linspace(0.9, 1.8, 4), 4, 2
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, R, D)
    assert gpu.sys_words
    lat = oracle.Lattice(shape)
    J = couplings("gaussian", shape, 3, D, 7).reshape((D, -1))
    rng = np.random.default_rng(3)
    S = R * len(temps)
    spins = (2 * rng.integers(0, 2, size=(D, S, lat.n_spins)) - 1).astype(np.int8)
    for d in range(D):
        gpu.set_spins(spins[d], d)
    for d in range(D):
        assert np.array_equal(gpu.get_spins(d).reshape(S, -1), spins[d])
    e_g, m_g = gpu.op_energies_mags()
    for d in range(D):
        e_c, m_c = lat.energies_mags(spins[d], J[d])
        np.testing.assert_allclose(e_g[d], e_c, rtol=1e-5, atol=1e-6)
        assert np.array_equal(m_g[d], m_c)
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, R, D)
    kw = dict(warmup_ratio=0.25, overlap_cluster_update_interval=2, overlap_cluster_mode="wolff")
    rg = gpu.sample(9, "metropolis", exact_log=True, **kw)
    rc = cpu.sample(9, "metropolis", **kw)
    assert_state_equal(gpu, cpu, D)
    for k in ("mags2", "overlap2", "link_overlap"):
        assert np.array_equal(rg[k], rc[k]), k


def test_sys_word_device_log_agrees_statistically(oracle):
    # production mode (hardware log2): trajectories may differ from the host-libm tables in rare last-ulp cases, observables must not
    shape, temps = (4, 4, 32), np.linspace(0.9, 1.8, 4)
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, 4, 4)
    assert gpu.sys_words
    for mode in ("metropolis", "gibbs"):
        rg = gpu.sample(400, mode, exact_log=False, pt_interval=1)
        rc = cpu.sample(400, mode, pt_interval=1)
        np.testing.assert_allclose(rg["energies"], rc["energies"], atol=0.02)
        np.testing.assert_allclose(rg["overlap2"], rc["overlap2"], atol=0.05)


def test_sys_word_autocorrelation_and_equilibration_diagnostic(oracle):
    """The per-sweep consumers of magnetisations and pair dots on a system-word handle (integer inputs: exact) and of the f32
    energies (tolerance); packed rows in cluster mode take the same options (the fold stays inside the kernel, the pushes read
    what it left in global memory)."""
    shape, temps, R, D = (4, 4, 32), np.linspace(0.9, 1.8, 4), 4, 2
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, R, D)
    assert gpu.sys_words
    kw = dict(warmup_ratio=0.25, autocorrelation_max_lag=4, equilibration_diagnostic=True)
    rg = gpu.sample(140, "metropolis", exact_log=True, **kw)
    rc = cpu.sample(140, "metropolis", **kw)
    assert_state_equal(gpu, cpu, D)
    for k in ("mags2", "overlap2", "link_overlap", "mags2_tau", "overlap2_tau", "equil_sweeps", "equil_link_overlap_avg"):
        assert np.array_equal(rg[k], rc[k]), k
    np.testing.assert_allclose(rg["equil_energy_avg"], rc["equil_energy_avg"], rtol=1e-5, atol=1e-7)
    # ferromagnet on packed rows, clusters of R CTAs: autocorrelation pushes after a fused recorded sweep
    gpu, cpu = make_pair(oracle, (4, 64), "ferro", [1.8, 2.27, 2.8], 2, 1)
    assert gpu.rows_packed
    for kw in (dict(autocorrelation_max_lag=5), dict(autocorrelation_max_lag=3, equilibration_diagnostic=True, pt_interval=2)):
        rg = gpu.sample(150, "gibbs", **kw)
        rc = cpu.sample(150, "gibbs", **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)


def gpu_couplings(shape, offsets, D):
    z = len(shape) if offsets is None else len(offsets)
    J = couplings("gaussian", shape, z, D, 7)
    return J.reshape((D, -1))


def test_gaussian_device_logf_agrees_statistically(oracle):
    # production mode (device logf): trajectories may differ in rare last-ulp cases, observables must not
    shape, temps = (6, 6, 6), np.linspace(0.9, 1.8, 4)
    gpu, cpu = make_pair(oracle, shape, "gaussian", temps, 2, 4)
    rg = gpu.sample(400, "metropolis", exact_log=False)
    rc = cpu.sample(400, "metropolis")
    np.testing.assert_allclose(rg["energies"], rc["energies"], atol=0.02)


def test_operator_entry_points_match_oracle(oracle):
    shape, temps, R, D = (6, 6, 6), [0.9, 1.2, 1.5], 4, 3
    for layout, kind, DD in (("int8", "bimodal", D), ("int8", "gaussian", D), ("msc", "bimodal", 37)):
        gpu, cpu = make_pair(oracle, shape, kind, temps, R, DD, layout=layout)
        lat = oracle.Lattice(shape)
        J = couplings(kind, shape, 3, DD, 7)
        colour = __import__("peapods_b200").colouring(shape)[0]
        rng = np.random.default_rng(5)
        S, N = R * len(temps), lat.n_spins
        spins = (2 * rng.integers(0, 2, size=(DD, S, N)) - 1).astype(np.int8)
        ids = np.stack([np.concatenate([r * len(temps) + rng.permutation(len(temps)) for r in range(R)]) for _ in range(DD)])
        for d in range(DD):
            gpu.set_system_ids(ids[d], d)
            gpu.set_spins(spins[d], d)
            assert np.array_equal(gpu.get_spins(d), spins[d].reshape(-1))
            assert np.array_equal(gpu.get_system_ids(d), ids[d])
        # energies / magnetisations (spins/energy.rs:59-76)
        e_g, m_g = gpu.op_energies_mags()
        for d in range(DD):
            e_c, m_c = lat.energies_mags(spins[d], J[d])
            assert np.array_equal(m_g[d], m_c)
            if kind == "gaussian":
                np.testing.assert_allclose(e_g[d], e_c, rtol=1e-5, atol=1e-6)
            else:
                assert np.array_equal(e_g[d], e_c)
        # overlap dots (statistics/overlap.rs:259-281)
        ds, dl = gpu.op_overlap()
        T = len(temps)
        for d in range(DD):
            for p in range(R // 2):
                for t in range(T):
                    a, b = ids[d][(2 * p) * T + t], ids[d][(2 * p + 1) * T + t]
                    assert (ds[d, p, t], dl[d, p, t]) == lat.overlap_dots(spins[d][a], spins[d][b])
        # one sweep with a caller-chosen sweep index (mcmc/sweep.rs:220-284)
        temps_full = np.tile(np.asarray(temps, np.float32), R)
        for sweep_mode, code in (("metropolis", oracle.SWEEP_METROPOLIS), ("gibbs", oracle.SWEEP_GIBBS)):
            gpu.op_sweep(sweep_mode, 77, exact_log=True)
            for d in range(DD):
                if layout == "msc":
                    key = oracle.lib().orc_splitmix64(1234 ^ oracle.lib().orc_splitmix64(0x6D73635F67726F75 ^ (d >> 5)))
                else:
                    key = oracle.lib().orc_realization_seed(1234, d)
                lat.sweep_philox(spins[d], J[d], temps_full, ids[d], colour, key, 77, code, use_lookup=True,
                                 stream_is_slot=(layout == "msc"))
                assert np.array_equal(gpu.get_spins(d), spins[d].reshape(-1)), (layout, kind, sweep_mode, d)


# ---- slab layout: one 3-D ferromagnet with stride geometry, cut along x0 (single-GPU emulation of the ranks) ----
SLAB_CASES = [
    # shape, temps, ranks
    ((4, 4, 8), [3.5, 4.51, 5.5], 1),
    ((8, 6, 16), [4.0, 4.51], 2),
    ((8, 6, 16), [4.0, 4.51], 4),
    ((2, 2, 8), [4.51], 1),           # L = 2 along x0 and x1: fwd == bwd neighbour, both counted
    ((12, 4, 24), [2.0, 4.51, 9.0], 3),
    # shape[2] % 64 == 0: one bit per spin (pp_kernels_slabp.cuh), packed draw mapping, bond counts inside the colour-1 pass
    ((4, 4, 64), [3.5, 4.51, 5.5], 1),
    ((8, 6, 128), [4.0, 4.51], 2),
    ((8, 6, 64), [4.0, 4.51], 4),
    ((2, 2, 64), [4.51], 1),
    ((12, 4, 192), [2.0, 4.51, 9.0, 1e9], 3),   # T = 1e9: a threshold at 2^24 takes the 7-threshold form of the kernel
]


@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
@pytest.mark.parametrize("shape,temps,ranks", SLAB_CASES)
def test_slab_layout_is_bit_exact(oracle, shape, temps, ranks, mode):
    import peapods_b200 as pb

    temps = np.asarray(temps, np.float32)
    colour, n_colours = pb.colouring(shape)
    assert n_colours == 2
    gpu = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 4321, layout="slab", slab_ranks=ranks, slab_rank=-1)
    assert gpu.layout == "slab" and gpu.n_local_spins == int(np.prod(shape))
    assert gpu.slab_packed == (shape[2] % 64 == 0)
    J = np.ones(tuple(shape) + (3,), np.float32)
    cpu = oracle.Sim(shape, J, temps, n_replicas=1, seed=4321, colour=colour,
                     rng_mode=oracle.RNG_PHILOX_PACKED if gpu.slab_packed else oracle.RNG_PHILOX)
    assert_state_equal(gpu, cpu, 1)
    for n_sweeps, interval in ((1, None), (2, None), (17, 1), (40, 3)):
        kw = dict(warmup_ratio=0.25, pt_interval=interval, pt_schedule="full_ladder")
        rg = gpu.sample(n_sweeps, mode, **kw)
        rc = cpu.sample(n_sweeps, mode, **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)
        assert np.array_equal(gpu.get_energies(0), cpu.energies(0))
    gpu.reset()
    cpu.reset()
    assert_state_equal(gpu, cpu, 1)


@pytest.mark.parametrize("shape", [(4, 6, 8), (4, 6, 64)])
def test_slab_set_spins_and_operator_entry_points(oracle, shape):
    import peapods_b200 as pb

    temps = np.asarray([4.51], np.float32)
    gpu = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 5, layout="slab", slab_ranks=2, slab_rank=-1)
    rng = np.random.default_rng(0)
    spins = (2 * rng.integers(0, 2, size=int(np.prod(shape))) - 1).astype(np.int8)
    gpu.set_spins(spins)
    assert np.array_equal(gpu.get_spins(), spins)
    e, m = gpu.op_energies_mags()
    s3 = spins.reshape(shape).astype(np.int64)
    bonds = sum(int((s3 * np.roll(s3, -1, axis=a)).sum()) for a in range(3))
    assert m[0, 0] == int(s3.sum())
    assert e[0, 0] == np.float32(bonds) / np.float32(s3.size)


def test_slab_rejects_what_it_cannot_hold():
    import peapods_b200 as pb

    t = np.asarray([4.5], np.float32)
    with pytest.raises(ValueError):
        pb.IsingSimulation([4, 4, 6], "ferro", t, 1, None, 1, layout="slab")          # shape[2] % 8
    with pytest.raises(ValueError):
        pb.IsingSimulation([4, 4, 8], "ferro", t, 2, None, 1, layout="slab")          # replicas
    with pytest.raises(ValueError):
        pb.IsingSimulation([6, 4, 8], "ferro", t, 1, None, 1, layout="slab", slab_ranks=2, slab_rank=-1)  # 6 % 4
    with pytest.raises(ValueError):
        pb.IsingSimulation([4, 4, 8], np.ones((4, 4, 8, 3), np.float32), t, 1, None, 1, layout="slab")


# ---- per-row stride tables (pp_kernels_rows.cuh): shapes whose last extent is a multiple of 8 take that path; it must
# be bit-identical to the table-driven kernels (PP_NO_ROWS=1) and to the oracle ----
ROWS_CASES = [
    # shape, kind, offsets, temps, R, D
    ((4, 16), "ferro", TRI, [3.2, 3.64, 4.1], 2, 1),                   # 4-colour triangular, shifted neighbour rows
    ((8, 8), "bimodal", TRI, [1.0, 3.64], 2, 2),
    ((6, 4, 8), "gaussian", None, [0.9, 1.4], 2, 2),
    ((4, 4, 16), "diluted", None, [0.8, 1.7], 3, 2),
    ((2, 24), "bimodal", None, [1.1, 2.2], 2, 1),                      # L=2 rows: forward == backward neighbour row
    ((4, 4, 8), "bimodal", FCC, [2.0, 9.8], 2, 1),
    ((4, 2, 2, 8), "gaussian", None, [1.5, 6.0], 2, 1),                # 4-D
]


@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", ROWS_CASES)
@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
@pytest.mark.parametrize("resident", ["1", "0"])  # one CTA per realization, many sweeps per launch / one launch per colour pass
def test_row_table_kernels_are_bit_exact(oracle, monkeypatch, shape, kind, offsets, temps, R, D, mode, resident):
    monkeypatch.setenv("PP_RESIDENT", resident)
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    assert_state_equal(gpu, cpu, D)
    gaussian = kind == "gaussian"  # f32 energies are tolerance-checked, so PT (which branches on them) stays off there
    for n_sweeps, interval in ((1, None), (19, 1), (30, 4)):
        interval = None if gaussian else interval
        kw = dict(warmup_ratio=0.25, pt_interval=interval, pt_schedule="full_ladder")
        rg = gpu.sample(n_sweeps, mode, exact_log=True, **kw)
        rc = cpu.sample(n_sweeps, mode, **kw)
        assert_state_equal(gpu, cpu, D)
        if not gaussian:
            assert_results_equal(rg, rc)
            continue
        for k in ("mags", "mags2", "mags4", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
            assert np.array_equal(rg[k], rc[k]), k
        np.testing.assert_allclose(rg["energies"], rc["energies"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(rg["energies2"], rc["energies2"], rtol=2e-5, atol=1e-7)


# ---- ferromagnets with rows of 64 k sites: one bit per spin, a system resident in shared memory per launch (pp_kernels_prows.cuh),
# packed draw mapping; the small cases switch the one-CTA-per-realization kernel off, which would otherwise take them ----
PROWS_CASES = [
    # shape, offsets, temps, R
    ((4, 64), None, [1.5, 2.27, 3.0, 1e9], 2),             # T = 1e9: a count at 2^24 takes the all-thresholds form
    ((8, 128), TRI, [3.0, 3.64, 4.2], 2),                   # the geometry of BASELINE config 3 (4 colours, shifted neighbour rows)
    ((4, 4, 64), None, [4.0, 4.51, 5.0], 2),
    ((2, 64), None, [2.0, 2.5], 3),                         # two rows: fwd == bwd neighbour row; odd replica count
    ((4, 2, 2, 64), None, [6.0, 6.7], 2),                   # 4-D: four forward directions
]


@pytest.mark.parametrize("shape,offsets,temps,R", PROWS_CASES)
@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
def test_packed_row_kernels_are_bit_exact(oracle, monkeypatch, shape, offsets, temps, R, mode):
    monkeypatch.setenv("PP_RESIDENT", "0")
    gpu, cpu = make_pair(oracle, shape, "ferro", temps, R, 1, offsets)
    assert gpu.rows_packed
    assert_state_equal(gpu, cpu, 1)
    seen = []
    for n_sweeps, interval, schedule in ((1, None, "single_random_edge"), (19, 1, "full_ladder"), (70, 4, "single_random_edge"), (90, None, "full_ladder")):
        kw = dict(warmup_ratio=0.25, pt_interval=interval, pt_schedule=schedule)
        rg = gpu.sample(n_sweeps, mode, on_sweep=seen.append, **kw)
        rc = cpu.sample(n_sweeps, mode, **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)
        assert np.array_equal(gpu.get_energies(0), cpu.energies(0))
    assert len(seen) == 180
    # operator entry points and a caller-written configuration go through the packed words too
    rng = np.random.default_rng(4)
    spins = (2 * rng.integers(0, 2, size=gpu.n_spins * R * len(temps)) - 1).astype(np.int8)
    gpu.set_spins(spins, 0)
    assert np.array_equal(gpu.get_spins(0), spins)
    lat = oracle.Lattice(shape, offsets)
    z = len(shape) if offsets is None else len(offsets)
    e_c, m_c = lat.energies_mags(spins.reshape(R * len(temps), -1), np.ones(tuple(shape) + (z,), np.float32))
    e_g, m_g = gpu.op_energies_mags()
    assert np.array_equal(e_g[0], e_c) and np.array_equal(m_g[0], m_c)
    gpu.reset()
    cpu.reset()
    assert_state_equal(gpu, cpu, 1)


CLUSTER_RESIDENT_CASES = [
    # shape, temps, R: small ferromagnets with rows of 32 k sites -> one bit per spin, the systems of a realization spread over a
    # thread-block cluster (prows_cluster_resident_kernel) whenever they divide evenly over 8 / 4 / 2 CTAs
    ((4, 32), np.linspace(1.8, 2.8, 4), 2),            # 8 systems of 4 words: one system per CTA, warps that span no whole system
    ((4, 4, 32), np.linspace(4.0, 5.0, 3), 2),         # three forward directions; 6 systems: a cluster of 2
    ((8, 64), np.linspace(1.8, 2.8, 4), 4),            # rows of two words, two replica pairs per temperature
    ((16, 32), [2.0, 2.27, 2.6, 3.0, 1e9], 2),         # 10 systems: a cluster of 2; T = 1e9: the all-thresholds form
    ((32, 32), np.linspace(1.5, 3.0, 5), 1),           # 5 systems: no even split, the one-CTA form; a single replica: no pairs
    ((32, 32), np.linspace(1.5, 3.0, 8), 1),           # 8 systems on 8 CTAs without replica pairs
]


@pytest.mark.parametrize("shape,temps,R", CLUSTER_RESIDENT_CASES)
@pytest.mark.parametrize("mode", ["metropolis", "gibbs"])
def test_cluster_resident_kernel_shapes_are_bit_exact(oracle, shape, temps, R, mode):
    gpu, cpu = make_pair(oracle, shape, "ferro", temps, R, 1)
    assert gpu.rows_packed
    assert_state_equal(gpu, cpu, 1)
    for n_sweeps, interval, schedule in ((1, None, "single_random_edge"), (23, 1, "full_ladder"), (61, 1, "single_random_edge"),
                                         (40, 3, "single_random_edge"), (300, 2, "full_ladder")):
        kw = dict(warmup_ratio=0.25, pt_interval=interval, pt_schedule=schedule)
        rg = gpu.sample(n_sweeps, mode, **kw)
        rc = cpu.sample(n_sweeps, mode, **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)
        assert np.array_equal(gpu.get_energies(0), cpu.energies(0))


def test_cluster_resident_kernel_equals_the_one_cta_form(oracle, monkeypatch):
    import peapods_b200 as pb

    shape, temps, R = (32, 32), np.linspace(1.5, 3.0, 16).astype(np.float32), 2
    kw = dict(warmup_ratio=0.25, pt_interval=1)
    many = pb.IsingSimulation(list(shape), "ferro", temps, R, None, 5, layout="int8")
    monkeypatch.setenv("PP_NO_RESIDENT_CLUSTER", "1")
    one = pb.IsingSimulation(list(shape), "ferro", temps, R, None, 5, layout="int8")
    for n in (7, 400):
        ra, rb = many.sample(n, "metropolis", **kw), one.sample(n, "metropolis", **kw)
        assert np.array_equal(many.get_spins(0), one.get_spins(0)) and np.array_equal(many.get_system_ids(0), one.get_system_ids(0))
        assert_results_equal(ra, rb)


@pytest.mark.parametrize("cluster_mode", ["sw", "wolff"])
def test_packed_rows_run_cluster_moves_through_the_int8_view(oracle, monkeypatch, cluster_mode):
    monkeypatch.setenv("PP_RESIDENT", "0")
    gpu, cpu = make_pair(oracle, (8, 64), "ferro", [2.0, 2.27, 2.6], 2, 1)
    assert gpu.rows_packed
    for kw in (dict(cluster_update_interval=1, cluster_mode=cluster_mode, pt_interval=1),
               dict(cluster_update_interval=3, cluster_mode=cluster_mode, pt_interval=2),
               dict(overlap_cluster_update_interval=2, overlap_cluster_mode=cluster_mode, pt_interval=1)):
        rg = gpu.sample(25, "metropolis", **kw)
        rc = cpu.sample(25, "metropolis", **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)


@pytest.mark.parametrize("schedule", ["single_random_edge", "full_ladder"])
@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", [
    ((32, 32), "ferro", None, np.linspace(1.5, 3.0, 16), 2, 1),          # BASELINE config 1 (README quickstart)
    ((8, 8, 8), "bimodal", None, np.linspace(0.8, 1.6, 6), 4, 3),
    ((8, 16), "bimodal", TRI, [1.0, 1.5, 2.5], 3, 2),                    # unpaired third replica
])
def test_resident_kernel_long_runs_are_bit_exact(oracle, shape, kind, offsets, temps, R, D, schedule):
    """More sweeps than one resident launch holds (256), PT every sweep and every third sweep, on_sweep, reset replay."""
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    seen = []
    for n_sweeps, interval in ((300, 1), (259, 3)):
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule, on_sweep=seen.append)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, pt_schedule=schedule)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        for d in range(D):
            assert np.array_equal(gpu.get_energies(d), cpu.energies(d))
    assert seen == list(range(300)) + list(range(259))
    assert gpu.last_kernel_launches <= 3


# ---- integrated autocorrelation times (statistics/autocorrelation.rs driven from simulation/mod.rs:341-371,
# 551-594, 825-832): the device accumulators repeat the reference's f64 operations in its order -> bit-exact taus ----
@pytest.mark.parametrize("layout,shape,kind,offsets,temps,R,D", [
    ("int8", (8, 8), "ferro", None, np.linspace(1.8, 2.8, 4), 2, 1),
    ("int8", (4, 4, 4), "bimodal", None, [0.9, 1.3, 1.7], 4, 3),
    ("int8", (6, 6), "bimodal", TRI, [1.0, 2.5], 3, 2),                 # unpaired replica; table-driven kernels
    ("int8", (8, 8), "bimodal", None, [1.2, 2.0], 1, 2),                # one replica: no overlap2_tau
    ("msc", (8, 8, 8), "bimodal", None, np.linspace(0.8, 1.6, 3), 2, 40),
    ("msc", (16, 16, 16), "bimodal", None, [0.9, 1.3], 4, 33),          # msc3d, in-sweep energy path, chunked streams
])
def test_autocorrelation_taus_are_bit_exact(oracle, layout, shape, kind, offsets, temps, R, D):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets, layout=layout)
    for n_sweeps, lag, interval in ((64, 5, 1), (41, 100, None), (3, 4, 2)):  # ring wraps; lag clamped to recorded / 4; to 1
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, autocorrelation_max_lag=lag)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, autocorrelation_max_lag=lag)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        assert ("overlap2_tau" in rg) == (R >= 2) and "mags2_tau" in rg
        assert np.array_equal(gpu.last_per_sample_taus[:, : (2 if R >= 2 else 1)], cpu.last_per_sample_taus[:, : (2 if R >= 2 else 1)])


# ---- equilibration diagnostic (statistics/equilibration.rs driven from simulation/mod.rs:511-541) ----
@pytest.mark.parametrize("layout,shape,kind,offsets,temps,R,D", [
    ("int8", (8, 8), "bimodal", None, [1.0, 1.8, 2.6], 2, 2),
    ("int8", (6, 6), "bimodal", TRI, [1.0, 2.5], 3, 1),                 # table-driven kernels, unpaired replica
    ("int8", (8, 8), "ferro", None, [2.0, 2.6], 1, 1),                  # one replica: link-overlap rows stay zero
    ("msc", (8, 8, 8), "bimodal", None, np.linspace(0.8, 1.6, 3), 2, 40),
    ("msc", (16, 16, 16), "bimodal", None, [0.9, 1.3], 4, 33),          # msc3d: overlap dots without the fold during warm-up
])
def test_equilibration_diagnostic_is_bit_exact(oracle, layout, shape, kind, offsets, temps, R, D):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets, layout=layout)
    for n_sweeps, interval in ((300, 3), (128, None), (20, 1)):  # checkpoints 128, 256, 300 / 128 / 20
        rg = gpu.sample(n_sweeps, "metropolis", pt_interval=interval, equilibration_diagnostic=True, autocorrelation_max_lag=3)
        rc = cpu.sample(n_sweeps, "metropolis", pt_interval=interval, equilibration_diagnostic=True, autocorrelation_max_lag=3)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        assert rg["equil_sweeps"].dtype == np.uint64 and rg["equil_energy_avg"].shape == (len(rg["equil_sweeps"]), len(temps))
        assert np.array_equal(gpu.last_per_sample_equil, cpu.last_per_sample_equil)


def test_equilibration_delta_matches_the_reference_formula():
    """spin_models.py:322-341 on a synthetic result (no sampling needed beyond the keys)."""
    from peapods_b200 import Ising

    model = Ising((4, 4), couplings="bimodal", temperatures=np.array([1.0, 2.0]), n_replicas=2, seed=9)
    model.sample(130, equilibration_diagnostic=True)
    sweeps, delta = model.equilibration_delta()
    assert list(sweeps) == [128, 130]
    expect = model._equil_energy_avg - (1.0 / model.temperatures) * 2 * (1 - model._equil_link_overlap_avg)
    assert np.array_equal(delta, expect)


# ---- Fortuin-Kasteleyn cluster updates (clusters/fk.rs:28-171 driven from simulation/mod.rs:434-470).  Bond and
# cluster draws are counter-based and clusters are named by their smallest site, so label propagation on the GPU and union-find
# in the oracle must produce the same configurations ----
@pytest.mark.parametrize("cluster_mode", ["sw", "wolff"])
@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", [
    ((16, 16), "ferro", None, np.linspace(1.8, 2.8, 5), 2, 1),          # row-table kernels, labels in shared memory
    ((8, 8, 8), "bimodal", None, [0.9, 1.5, 2.4], 2, 2),
    ((6, 6), "diluted", TRI, [1.0, 2.5], 3, 2),                         # zeros in the couplings, table-driven kernels
    ((5, 7), "bimodal", None, [0.8, 2.5], 2, 1),                        # odd extents
    ((256, 256), "ferro", None, [2.1, 2.27, 2.5], 1, 1),                # labels in global scratch (64 Ki sites), critical clusters
])
def test_cluster_updates_are_bit_exact(oracle, cluster_mode, shape, kind, offsets, temps, R, D):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    big = int(np.prod(shape)) > 10000
    for n_sweeps, interval, pt in ((6, 1, None), (12, 1, 1), (9, 2, 3)) if big else ((40, 1, None), (30, 1, 1), (25, 3, 2)):
        kw = dict(cluster_update_interval=interval, cluster_mode=cluster_mode, pt_interval=pt)
        rg = gpu.sample(n_sweeps, "metropolis", **kw)
        rc = cpu.sample(n_sweeps, "metropolis", **kw)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)


def test_cluster_updates_need_the_int8_layout_and_unit_couplings(oracle):
    gpu, _ = make_pair(oracle, (4, 4, 8), "bimodal", [1.0, 2.0], 2, 32, layout="msc")
    before = gpu.get_spins(0).copy()
    with pytest.raises(ValueError, match="not implemented on the GPU sweep path"):
        gpu.sample(4, "metropolis", cluster_update_interval=1)
    assert np.array_equal(gpu.get_spins(0), before)
    gauss, _ = make_pair(oracle, (4, 4, 8), "gaussian", [1.0, 2.0], 2, 1)
    with pytest.raises(ValueError, match="not implemented on the GPU sweep path"):
        gauss.sample(4, "metropolis", cluster_update_interval=1)


# ---- Houdayer isoenergetic cluster move (clusters/overlap.rs:34-56, 146-339 from simulation/mod.rs:596-756) ----
@pytest.mark.parametrize("oc_mode", ["wolff", "sw"])
@pytest.mark.parametrize("shape,kind,offsets,temps,R,D", [
    ((8, 8, 8), "bimodal", None, [0.7, 1.0, 1.4], 2, 2),
    ((6, 6, 6), "bimodal", None, np.linspace(0.6, 1.6, 4), 4, 3),       # two pairs per temperature: shuffled pairing
    ((8, 8), "gaussian", TRI, [0.8, 1.6], 3, 2),                        # unpaired third replica; no couplings involved
    ((5, 7), "bimodal", None, [0.8, 2.5], 2, 1),
    ((64, 64), "bimodal", None, [0.6, 1.2], 2, 1),                      # 4096 sites: 2-D spin glass, large overlap clusters
    ((256, 256), "bimodal", None, [0.7, 1.3], 2, 1),                    # 65 536 sites: labels in global scratch
])
def test_houdayer_moves_are_bit_exact(oracle, oc_mode, shape, kind, offsets, temps, R, D):
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets)
    gaussian = kind == "gaussian"
    big = int(np.prod(shape)) > 10000
    for n_sweeps, interval, pt in ((5, 1, None), (6, 1, 1), (7, 3, 2)) if big else ((30, 1, None), (24, 1, 1), (25, 3, 2)):
        pt = None if gaussian else pt  # f32 energies are tolerance-checked: PT (which branches on them) stays off there
        kw = dict(overlap_cluster_update_interval=interval, overlap_cluster_mode=oc_mode, pt_interval=pt)
        rg = gpu.sample(n_sweeps, "metropolis", exact_log=True, **kw)
        rc = cpu.sample(n_sweeps, "metropolis", **kw)
        assert_state_equal(gpu, cpu, D)
        if not gaussian:
            assert_results_equal(rg, rc)


def test_houdayer_needs_two_replicas_and_the_int8_layout(oracle):
    one, _ = make_pair(oracle, (4, 4, 8), "bimodal", [1.0, 2.0], 1, 1)
    with pytest.raises(ValueError, match="overlap cluster requires n_replicas >= max group_size"):
        one.sample(4, "metropolis", overlap_cluster_update_interval=1)
    msc, _ = make_pair(oracle, (4, 4, 8), "bimodal", [1.0, 2.0], 2, 32, layout="msc")
    before = msc.get_spins(0).copy()
    with pytest.raises(ValueError, match="not implemented on the GPU sweep path"):
        msc.sample(4, "metropolis", overlap_cluster_update_interval=1, overlap_cluster_mode="sw")  # per-cluster coins need labels
    assert np.array_equal(msc.get_spins(0), before)


@pytest.mark.parametrize("shape,kind,offsets,D,R,temps", [
    ((8, 8), "bimodal", None, 40, 2, [0.7, 1.2, 2.0]),                  # generic multispin kernels, padded second word group
    ((6, 6), "bimodal", TRI, 33, 3, [1.0, 2.5]),                        # unpaired third ladder
    ((8, 8, 8), "bimodal", None, 45, 4, np.linspace(0.7, 1.5, 4)),      # msc3d, shuffled pairing of four ladders
    ((16, 16, 16), "bimodal", None, 64, 4, [0.8, 1.1, 1.4]),            # headline geometry: in-sweep energy path + chunks
])
def test_multispin_houdayer_moves_are_bit_exact(oracle, shape, kind, offsets, D, R, temps):
    """32 realizations per word: lane-uniform draws (pairing, seed scores) come from the group key, the clusters of the 32 lanes
    grow as one bit-parallel flood fill; the oracle runs the realizations one by one with union-find."""
    gpu, cpu = make_pair(oracle, shape, kind, temps, R, D, offsets, layout="msc")
    for n_sweeps, interval, pt in ((12, 1, None), (16, 1, 1), (13, 3, 2)):
        kw = dict(overlap_cluster_update_interval=interval, pt_interval=pt, pt_schedule="full_ladder" if pt == 2 else "single_random_edge")
        rg = gpu.sample(n_sweeps, "metropolis", **kw)
        rc = cpu.sample(n_sweeps, "metropolis", **kw)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)


def test_multispin_houdayer_with_few_active_sites(oracle):
    """Replicas that start identical and move at low temperature differ at a handful of sites: most lanes find no active site among the
    low-score candidates and take the full-scan fallback (or have no active site at all: the move is skipped for them)."""
    import ctypes as C

    shape, temps, R, D = (16, 16, 16), [0.45, 0.5], 2, 40
    gpu, cpu = make_pair(oracle, shape, "bimodal", temps, R, D, layout="msc")
    S, N = R * len(temps), int(np.prod(shape))
    for d in range(D):
        spins = gpu.get_spins(d).reshape(S, N).copy()
        spins[:] = spins[0]                                   # every system of the realization starts from the same configuration
        gpu.set_spins(spins.reshape(-1), d)
        ptr = oracle.lib().orc_sim_spins(cpu.h, d)
        np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int8)), shape=(S * N,))[:] = spins.reshape(-1)
    assert_state_equal(gpu, cpu, D)
    n_active = []
    for n_sweeps in (1, 3, 6):
        rg = gpu.sample(n_sweeps, "metropolis", overlap_cluster_update_interval=1, pt_interval=2)
        rc = cpu.sample(n_sweeps, "metropolis", overlap_cluster_update_interval=1, pt_interval=2)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        sp = gpu.get_spins(0).reshape(S, N)
        n_active.append(int(np.sum(sp[0] != sp[len(temps)])))
    assert 0 < max(n_active) < 200, n_active                    # the regime the test is about


def test_chunked_multistream_execution_of_the_widened_features_is_bit_exact(oracle, monkeypatch):
    """Autocorrelation pushes, equilibration sums and Houdayer moves run per chunk on the chunk's stream with offset views of their
    accumulators: five word groups in chunks of one or two over three streams must reproduce the unchunked oracle."""
    monkeypatch.setenv("PP_STREAMS", "3")
    monkeypatch.setenv("PP_CHUNK_GROUPS", "2")
    monkeypatch.setenv("PP_MACRO_BATCH", "4")
    shape, temps, R, D = (4, 4, 8), np.linspace(0.8, 1.6, 5), 4, 150   # 5 word groups, the last one padded
    gpu, cpu = make_pair(oracle, shape, "bimodal", temps, R, D, layout="msc")
    assert gpu.uses_msc3d
    for n_sweeps, interval in ((140, 1), (37, 3)):
        kw = dict(pt_interval=interval, autocorrelation_max_lag=6, equilibration_diagnostic=True, overlap_cluster_update_interval=2)
        rg = gpu.sample(n_sweeps, "metropolis", **kw)
        rc = cpu.sample(n_sweeps, "metropolis", **kw)
        assert_state_equal(gpu, cpu, D)
        assert_results_equal(rg, rc)
        assert np.array_equal(gpu.last_per_sample_taus, cpu.last_per_sample_taus)
        assert np.array_equal(gpu.last_per_sample_equil, cpu.last_per_sample_equil)


def test_headline_full_size_properties_and_oracle_checked_shard(oracle):
    """BASELINE configs[1] at its full size (16^3 +-J, 32 temperatures, 4 replicas, 4096 realizations: more than the oracle can
    replay in seconds), checked through size-independent properties, plus one shard the oracle does replay:
      * realizations are independent (simulation/mod.rs:887-903): the first 32 realizations of the full run equal a 32-realization
        handle bit for bit (per-realization means, spins, PT counters), and that handle equals the oracle;
      * every recorded sweep adds one count per (realization, temperature, pair) to the histogram (statistics/overlap.rs:300-310);
      * single_random_edge attempts exactly one edge per replica ladder and PT event (mcmc/tempering.rs:29-41);
      * the aggregated means are the in-order mean of the per-realization means (statistics/results.rs:165-180)."""
    import peapods_b200 as pb

    shape, T, R, D, n_sweeps, sub = (16, 16, 16), 32, 4, 4096, 12, 32
    temps = np.linspace(0.8, 1.4, T).astype(np.float32)
    rng = np.random.default_rng(2026)
    J = (2 * rng.integers(0, 2, size=(D,) + shape + (3,), dtype=np.int8) - 1).astype(np.float32)
    kw = dict(pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25, per_sample=False)
    full = pb.IsingSimulation(list(shape), J, temps, R, None, 4242, layout="msc")
    res = full.sample(n_sweeps, "metropolis", **kw)
    means = full.last_per_sample_means.copy()  # [D, 11, T]
    n_rec, P = n_sweeps - 3, R // 2

    hist = np.stack(res["overlap_histogram"])
    assert np.array_equal(hist.sum(axis=1), np.full(T, D * P * n_rec, np.uint64))
    pt = res["per_disorder"]["parallel_tempering"]
    assert np.array_equal(pt["edge_attempts"].sum(axis=1), np.full(D, R * n_sweeps, np.uint64))
    assert np.all(pt["edge_acceptances"] <= pt["edge_attempts"])
    keys = ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2",
            "link_overlap4")
    for k, key in enumerate(keys):
        acc = np.zeros(T)
        for d in range(D):
            acc += means[d, k]
        assert np.array_equal(res[key], acc / D), key
    # energy.rs:92-109 sums + s_i s_j J per bond: satisfied bonds count positive, at most z' = 3 per site
    assert np.all(res["energies"] > 0.0) and np.all(res["energies"] < 3.0)
    assert np.all(np.diff(res["energies"]) < 0.0)  # fewer satisfied bonds at higher temperature (4096-realization average)

    part = pb.IsingSimulation(list(shape), J[:sub], temps, R, None, 4242, layout="msc")
    rp = part.sample(n_sweeps, "metropolis", **kw)
    assert np.array_equal(part.last_per_sample_means, means[:sub])
    for k in pt:
        assert np.array_equal(rp["per_disorder"]["parallel_tempering"][k], pt[k][:sub]), k
    for d in (0, 17, sub - 1):
        assert np.array_equal(part.get_spins(d), full.get_spins(d))
        assert np.array_equal(part.get_system_ids(d), full.get_system_ids(d))

    colour, _ = pb.colouring(shape, None)
    cpu = oracle.Sim(shape, J[:sub], temps, n_replicas=R, offsets=None, seed=4242, rng_mode=oracle.RNG_PHILOX_MSC, colour=colour)
    rc = cpu.sample(n_sweeps, "metropolis", pt_interval=1, pt_schedule="single_random_edge", warmup_ratio=0.25)
    for key in keys:
        assert np.array_equal(rp[key], rc[key]), key
    assert np.array_equal(np.stack(rp["overlap_histogram"]), np.asarray(rc["overlap_histogram"]))
    for d in (0, sub - 1):
        assert np.array_equal(part.get_spins(d), cpu.spins(d))


def test_config3_full_size_triangular_gibbs_is_bit_exact(oracle):
    """BASELINE configs[2] at its full size: 256 x 256 triangular ferromagnet (custom neighbor_offsets), Gibbs sweeps, 64
    temperatures around T_c = 4 / ln 3, 2 replicas — small enough for the oracle to replay a few sweeps outright."""
    tc = 4.0 / np.log(3.0)
    gpu, cpu = make_pair(oracle, (256, 256), "ferro", np.linspace(tc - 0.4, tc + 0.4, 64), 2, 1, TRI)
    assert_state_equal(gpu, cpu, 1)
    for n_sweeps, interval in ((3, None), (4, 1)):
        kw = dict(warmup_ratio=0.25, pt_interval=interval, pt_schedule="single_random_edge")
        rg = gpu.sample(n_sweeps, "gibbs", **kw)
        rc = cpu.sample(n_sweeps, "gibbs", **kw)
        assert_state_equal(gpu, cpu, 1)
        assert_results_equal(rg, rc)


def test_config4_full_size_gaussian_properties_and_oracle_checked_shard(oracle):
    """BASELINE configs[3] at its full size (32^3 Gaussian couplings, 48 temperatures, 4 replicas, 512 realizations) through
    size-independent properties, plus a 2-realization shard replayed by the oracle (spins bit-exact with exact_log, integer
    observables exact, f32 energies within 1e-5 relative — the tolerance BASELINE.json states)."""
    import peapods_b200 as pb

    shape, T, R, D, n_sweeps, sub = (32, 32, 32), 48, 4, 512, 4, 2
    temps = np.linspace(0.8, 1.8, T).astype(np.float32)
    J = np.random.default_rng(44).standard_normal((D,) + shape + (3,), dtype=np.float32)
    kw = dict(warmup_ratio=0.25, exact_log=True, per_sample=False)
    full = pb.IsingSimulation(list(shape), J, temps, R, None, 77)
    assert full.layout == "int8"
    res = full.sample(n_sweeps, "metropolis", **kw)
    means = full.last_per_sample_means.copy()
    n_rec, P = n_sweeps - 1, R // 2
    assert np.array_equal(np.stack(res["overlap_histogram"]).sum(axis=1), np.full(T, D * P * n_rec, np.uint64))
    keys = ("mags", "mags2", "mags4", "energies", "energies2", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2",
            "link_overlap4")
    for k, key in enumerate(keys):
        acc = np.zeros(T)
        for d in range(D):
            acc += means[d, k]
        assert np.array_equal(res[key], acc / D), key
    assert np.all(np.isfinite(res["energies"])) and np.all(res["energies"] > 0.0)
    assert np.all(np.abs(res["overlap"]) <= 1.0) and np.all(res["overlap2"] <= 1.0)

    part = pb.IsingSimulation(list(shape), J[:sub], temps, R, None, 77)
    rp = part.sample(n_sweeps, "metropolis", **kw)
    for d in range(sub):
        assert np.array_equal(part.get_spins(d), full.get_spins(d))
    exact_rows = [0, 1, 2, 5, 6, 7, 8, 9, 10]  # magnetisation and overlap moments come from integers
    assert np.array_equal(part.last_per_sample_means[:, exact_rows], means[:sub][:, exact_rows])
    # fp32-coupling energies: every bond-sum term is rounded to an integer number of 1 / escale on the spot, so the sums do not depend
    # on how the sites are grouped into threads and blocks (the two handles use 1 and 8 sites per thread)
    assert np.array_equal(part.last_per_sample_means[:, 3:5], means[:sub, 3:5])

    colour, _ = pb.colouring(shape, None)
    assert full.sys_words and part.sys_words  # the same site of 32 systems per word (pp_kernels_swords.cuh)
    cpu = oracle.Sim(shape, J[:sub], temps, n_replicas=R, offsets=None, seed=77, rng_mode=oracle.RNG_PHILOX_SYSQ, colour=colour)
    rc = cpu.sample(n_sweeps, "metropolis", warmup_ratio=0.25)
    assert_state_equal(part, cpu, sub)
    for k in ("mags", "mags2", "mags4", "overlap", "overlap2", "overlap4", "link_overlap", "link_overlap2", "link_overlap4"):
        assert np.array_equal(rp[k], rc[k]), k
    assert np.array_equal(np.stack(rp["overlap_histogram"]), np.asarray(rc["overlap_histogram"]))
    np.testing.assert_allclose(rp["energies"], rc["energies"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(rp["energies2"], rc["energies2"], rtol=2e-5, atol=1e-7)


def test_config5_full_size_slab_decomposition_properties(oracle):
    """BASELINE configs[4]: the 1024^3 ferromagnet at T_c on the slab layout.  The oracle replays a 256^3 lattice cut into 8 slabs;
    at the full size (2^30 sites) the checks are size-independent: the 8-slab decomposition (the ranks emulated on one GPU, halo
    planes exchanged per colour half-step) leaves the same 2^30 spins as the undivided lattice, and the device's integer bond /
    spin sums equal a host recount of the downloaded configuration (energy.rs:92-109 on integers)."""
    import peapods_b200 as pb

    temps = np.asarray([4.511], np.float32)
    shape = (256, 256, 256)
    colour, _ = pb.colouring(shape)
    gpu = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 11, layout="slab", slab_ranks=8, slab_rank=-1)
    assert gpu.slab_packed
    cpu = oracle.Sim(shape, np.ones(shape + (3,), np.float32), temps, n_replicas=1, seed=11, rng_mode=oracle.RNG_PHILOX_PACKED, colour=colour)
    rg, rc = gpu.sample(2, "metropolis", warmup_ratio=0.5), cpu.sample(2, "metropolis", warmup_ratio=0.5)
    assert_state_equal(gpu, cpu, 1)
    assert_results_equal(rg, rc)
    del gpu, cpu

    shape = (1024, 1024, 1024)
    whole = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 11, layout="slab", slab_ranks=1, slab_rank=-1)
    slabs = pb.IsingSimulation(list(shape), "ferro", temps, 1, None, 11, layout="slab", slab_ranks=8, slab_rank=-1)
    ra, rb = whole.sample(3, "metropolis", warmup_ratio=0.25), slabs.sample(3, "metropolis", warmup_ratio=0.25)
    for k in ("mags", "mags2", "mags4", "energies", "energies2"):
        assert np.array_equal(ra[k], rb[k]), k
    spins = slabs.get_spins()
    assert np.array_equal(whole.get_spins(), spins)
    e, m = slabs.op_energies_mags()
    s3 = spins.reshape(shape)
    assert m[0, 0] == int(s3.sum(dtype=np.int64))
    bonds = 0
    for a in range(3):
        bonds += int(np.multiply(s3, np.roll(s3, -1, axis=a), dtype=np.int8).sum(dtype=np.int64))
    assert e[0, 0] == np.float32(bonds) / np.float32(s3.size)
    assert 0.5 < float(e[0, 0]) < 3.0  # three sweeps from a random start at T_c: ordering has begun


def test_slab_layout_on_real_ranks_is_bit_exact():
    """The ncclSend / ncclRecv halo exchange and the ncclAllReduce of the bond counts on REAL ranks (one process per GPU,
    tools/slab_check.py under torchrun): spins, energies and result dicts of byte and bit-packed slabs equal the oracle's
    whole-lattice run.  Needs at least two visible GPUs; the single-GPU tests above emulate the ranks on one device."""
    import subprocess
    import sys

    import torch

    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("one GPU visible: real-rank slab exchange needs >= 2 (bench.py --gpus N runs its own N-rank check)")
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "slab_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SLAB_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_system_split_on_real_ranks_is_bit_exact():
    """the systems of ONE realization split over real ranks (tools/system_split_check.py under torchrun):
    energies / magnetisations / configurations all-gathered with NCCL, exchange decisions replayed on every rank; spins, system
    ids and result dicts equal the oracle's unsplit run on every rank.  Needs at least two visible GPUs."""
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the system split needs >= 2 (bench.py --gpus N runs its own N-rank check)")
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29534", str(ROOT / "tools" / "system_split_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SYSTEM_SPLIT_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_system_split_rejects_what_it_cannot_hold():
    import peapods_b200 as pb

    t = np.linspace(2.0, 3.0, 3).astype(np.float32)
    with pytest.raises(ValueError):   # 3 temperatures x 1 replica over 2 ranks
        pb.IsingSimulation([8, 8], "ferro", t, 1, None, 1, layout="int8", system_ranks=2, system_rank=0, nccl_unique_id=bytes(128))
    with pytest.raises(ValueError):   # several realizations: shard those by sample instead
        pb.IsingSimulation([8, 8], np.ones((2, 8, 8, 2), np.float32), t[:2], 1, None, 1, layout="int8", system_ranks=2, system_rank=0,
                           nccl_unique_id=bytes(128))
    with pytest.raises(ValueError):   # multispin layout
        pb.IsingSimulation([8, 8], "ferro", t[:2], 1, None, 1, layout="msc", system_ranks=2, system_rank=0, nccl_unique_id=bytes(128))


def test_generic_multispin_counters_hold_a_large_lattice():
    """Round-1 advisor finding: the table-driven multispin kernels count per thread in bit-sliced planes; with 12 planes a thread
    overflowed once ceil(N / 256) * z passed 4095 (2-D 1024^2: 8192 adds per thread).  20 planes now, guarded in pp_create:
    energies, magnetisations and overlap dots of random configurations on 1024 x 1024 (N z = 2^21) against a host count."""
    import peapods_b200 as pb

    shape, D, R = (1024, 1024), 32, 2
    rng = np.random.default_rng(21)
    J = (2 * rng.integers(0, 2, size=(D,) + shape + (2,), dtype=np.int8) - 1).astype(np.float32)
    gpu = pb.IsingSimulation(list(shape), J, np.asarray([2.0], np.float32), R, None, 5, layout="msc")
    assert gpu.layout == "msc" and not gpu.uses_msc3d
    N = shape[0] * shape[1]
    check = (0, 17, 31)
    spins = {}
    for d in check:   # the other lanes keep their INIT-domain configuration
        spins[d] = (2 * rng.integers(0, 2, size=(R, N), dtype=np.int8) - 1).astype(np.int8)
        gpu.set_spins(spins[d], d)
    e, m = gpu.op_energies_mags()
    ds, dl = gpu.op_overlap()
    for d in check:
        s = spins[d].reshape((R,) + shape).astype(np.int64)
        Jd = J[d].astype(np.int64)
        for r in range(R):
            bonds = int((s[r] * np.roll(s[r], -1, axis=0) * Jd[..., 0]).sum() + (s[r] * np.roll(s[r], -1, axis=1) * Jd[..., 1]).sum())
            assert m[d, r] == int(s[r].sum())
            assert e[d, r] == np.float32(bonds) / np.float32(N)
        q = s[0] * s[1]
        assert ds[d, 0, 0] == int(q.sum())
        assert dl[d, 0, 0] == int((q * np.roll(q, -1, axis=0)).sum() + (q * np.roll(q, -1, axis=1)).sum())


def test_slice_signature_operators_match_the_oracle(oracle):
    """SURVEY 8b "Granularity": the reference's operator functions on HOST slices (peapods_b200.ops over pp_slice_*): sweep,
    energies + magnetisations, overlap dots against the oracle's operators; exchange events against the handle-based operator."""
    import peapods_b200 as pb
    from peapods_b200 import ops

    shape, temps, R = (6, 6, 6), np.asarray([0.9, 1.2, 1.5], np.float32), 4
    T, S = len(temps), R * len(temps)
    lat = oracle.Lattice(shape)
    N = lat.n_spins
    colour = pb.colouring(shape)[0]
    rng = np.random.default_rng(9)
    ids = np.concatenate([r * T + rng.permutation(T) for r in range(R)]).astype(np.int64)
    for kind in ("bimodal", "gaussian"):
        J = np.asarray(couplings(kind, shape, 3, 1, 3), np.float32).reshape(tuple(shape) + (3,))
        spins = (2 * rng.integers(0, 2, size=(S, N)) - 1).astype(np.int8)
        e_c, m_c = lat.energies_mags(spins, J)
        e_g, m_g = ops.compute_energies_and_magnetizations(shape, spins, J, S)
        assert np.array_equal(m_g, m_c)
        if kind == "gaussian":
            np.testing.assert_allclose(e_g, e_c, rtol=1e-5, atol=1e-6)
        else:
            assert np.array_equal(e_g, e_c)
        ds, dl = ops.overlap_dots(shape, spins, temps, ids, n_replicas=R)
        for p in range(R // 2):
            for t in range(T):
                assert (ds[p, t], dl[p, t]) == lat.overlap_dots(spins[ids[(2 * p) * T + t]], spins[ids[(2 * p + 1) * T + t]])
        for fn, code in ((ops.metropolis_sweep, oracle.SWEEP_METROPOLIS), (ops.gibbs_sweep, oracle.SWEEP_GIBBS)):
            mine, ref = spins.copy(), spins.copy()
            fn(shape, mine, J, temps, ids, n_replicas=R, seed=1234, sweep_index=77)
            lat.sweep_philox(ref, J, np.tile(temps, R), ids, colour, oracle.lib().orc_realization_seed(1234, 0), 77, code, use_lookup=True)
            assert np.array_equal(mine, ref), (kind, code)
    # exchange events: the slice form equals the handle form on the same energies
    J = np.asarray(couplings("bimodal", shape, 3, 1, 3), np.float32).reshape(tuple(shape) + (3,))
    spins = (2 * rng.integers(0, 2, size=(S, N)) - 1).astype(np.int8)
    e_g, _ = ops.compute_energies_and_magnetizations(shape, spins, J, S)
    for schedule, fn in (("single_random_edge", ops.parallel_tempering), ("full_ladder", ops.parallel_tempering_full_ladder)):
        h = pb.IsingSimulation(list(shape), J, temps, R, None, 1234, layout="int8")
        h.set_system_ids(ids, 0)
        h.set_spins(spins, 0)
        h.op_energies_mags()
        cur = ids.copy()
        for event in range(6):
            h.op_pt(schedule, event)
            kw = dict(first_parity=event % 2) if schedule == "full_ladder" else {}
            cur = fn(shape, e_g, temps, cur, n_replicas=R, seed=1234, pt_event=event, **kw)
            assert np.array_equal(cur, h.get_system_ids(0)), (schedule, event)
        assert sorted(cur.tolist()) == list(range(S)) and not np.array_equal(cur, ids)
