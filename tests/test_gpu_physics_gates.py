"""The reference's own physics gates, run through the `Ising` mirror on the GPU engine with the reference's
recipes, sizes, sweep counts and tolerances:

  tests/binder_crossings.py + tests/utils.py:39-47    Binder cumulants of L = 8, 16, 32 cross at T_c within 0.05 (2-D square, triangular)
  tests/spin_glass_crossings.py:14-51, utils.py:15-36 3-D EA +-J, L = 8, 10: histogram SG-Binder == direct SG-Binder within 0.05 at every
                                                      temperature; curves of the two sizes within 0.3 at T_c = 1.102
  tests/overlap_histogram.py:12-101                   3-D Gaussian 8^3 at T = 1.4: |<q>| < 0.1, P(q) symmetric within 0.25, thermalisation
                                                      Delta < 0.15, A(q) small, I(q)/X(q) within 0.15 of 1
and the north-star's free-running bar: <e>, SG-Binder and P(q) of the GPU engine within 2 sigma of the reference-faithful CPU
restatement (typewriter order, xoshiro streams) over independent seeds, bin by bin."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TC_SQUARE = 2.0 / np.log(1 + np.sqrt(2))
TC_TRIANGULAR = 4.0 / np.log(3)
TC_EA_3D = 1.102
SEED = 42  # tests/binder_crossings.py:19


def crossing_spread(temps, curves, tc):
    at_tc = [np.interp(tc, temps, c) for c in curves]  # tests/utils.py:41
    return max(at_tc) - min(at_tc), at_tc


@pytest.mark.parametrize("name,tc,half_width,geometry", [("square", TC_SQUARE, 0.3, None), ("triangular", TC_TRIANGULAR, 0.4, "tri")])
def test_binder_cumulants_cross_at_tc(name, tc, half_width, geometry):
    import peapods_b200 as pb

    temps = np.linspace(tc - half_width, tc + half_width, 32).astype(np.float32)
    curves = []
    for L in (8, 16, 32):
        model = pb.Ising((L, L), temperatures=temps, n_replicas=2, seed=SEED, geometry=geometry)
        model.sample(10000, sweep_mode="metropolis", cluster_update_interval=1, cluster_mode="sw", pt_interval=1, warmup_ratio=0.25)
        curves.append(model.binder_cumulant)
    spread, at_tc = crossing_spread(temps, curves, tc)
    assert spread < 0.05, (name, at_tc)
    assert all(0.55 < b < 0.66 for b in at_tc), at_tc  # the universal value of the periodic 2-D Ising class is 0.61
    for c in curves:   # ordered at the cold end (-> 2/3), disordering at the hot end; larger lattices fall faster
        assert c[0] > 0.6 and c[-1] < c[0]
    assert curves[2][-1] < curves[0][-1]


def histogram_binder(model):
    n_bins = model.n_spins + 1
    q = np.linspace(-1, 1, n_bins)
    out = []
    for t in range(model.n_temps):  # tests/utils.py:21-31
        p = model.overlap_histogram[t].astype(np.float64)
        p /= p.sum()
        q2, q4 = (q ** 2 * p).sum(), (q ** 4 * p).sum()
        out.append(1 - q4 / (3 * q2 ** 2))
    return np.array(out)


def test_spin_glass_histogram_binder_equals_direct_and_sizes_cross():
    import peapods_b200 as pb

    temps = np.linspace(0.8, 1.4, 12).astype(np.float32)
    curves = []
    for L in (8, 10):
        model = pb.Ising((L, L, L), couplings="bimodal", temperatures=temps, n_replicas=2, n_disorder=25, seed=SEED)
        model.sample(10000, sweep_mode="metropolis", pt_interval=1, overlap_cluster_update_interval=1, warmup_ratio=0.25)
        err = np.abs(histogram_binder(model) - model.sg_binder)
        assert err.max() < 0.05, (L, err)
        assert np.all(np.diff(model.sg_binder) < 0.05)  # the SG Binder ratio falls with temperature
        curves.append(model.sg_binder)
    spread, at_tc = crossing_spread(temps, curves, TC_EA_3D)
    assert spread < 0.3, at_tc


def test_gaussian_spin_glass_overlap_histogram_checks():
    import peapods_b200 as pb
    from peapods_b200.sweep import cumulative_overlap_ratio

    D = 64
    model = pb.Ising((8, 8, 8), couplings="gaussian", temperatures=np.array([1.4], dtype=np.float32), n_replicas=2, n_disorder=D, seed=SEED)
    model.sample(40000, sweep_mode="metropolis", pt_interval=1, overlap_cluster_update_interval=1, warmup_ratio=0.25,
                 equilibration_diagnostic=True)
    assert abs(model.overlap[0]) < 0.1
    hist = model.overlap_histogram[0].astype(float)
    assert hist.sum() == 30000 * D
    assert np.linalg.norm(hist - hist[::-1]) / np.linalg.norm(hist) < 0.25
    ps = model.per_sample_overlap_histogram
    assert ps.shape == (D, 1, 513) and model.per_sample_ql_at_q_sum.shape == (D, 1, 513)
    sweeps, delta = model.equilibration_delta(j_squared=1.0)
    assert abs(delta[-1, 0]) < 0.15, delta[:, 0]
    psf = ps.astype(float)
    mask = psf > 0
    mean_ql = np.where(mask, model.per_sample_ql_at_q_sum / np.where(mask, psf, 1), 0)
    a_s = np.where(mask, model.per_sample_ql2_at_q_sum / np.where(mask, psf, 1) - mean_ql ** 2, 0)
    denom = psf.sum(axis=0)
    a_q = np.where(denom > 0, (psf * a_s).sum(axis=0) / np.where(denom > 0, denom, 1), 0)
    a_mean = (a_q * denom).sum(axis=-1) / denom.sum(axis=-1)
    assert -1e-6 <= a_mean[0] < 0.05, a_mean
    q_grid, ratio, _, _ = cumulative_overlap_ratio(ps)
    assert np.max(np.abs(ratio[0, 1:len(q_grid) // 2] - 1.0)) < 0.15


def test_free_running_observables_agree_with_the_reference_faithful_cpu_run_within_two_sigma(oracle):
    """8^3 +-J, 4 temperatures, 2 replicas, Metropolis + PT: 16 GPU seeds (colour order, Philox) against 16 CPU seeds (typewriter
    order, one xoshiro stream per system).  <e>, <q^2>, SG-Binder per temperature and P(q) in 16 coarse bins per temperature."""
    import peapods_b200 as pb

    shape, R, n_sweeps, n_seeds = (8, 8, 8), 2, 6000, 16
    temps = np.asarray([0.9, 1.1, 1.4, 1.8], np.float32)
    rng = np.random.default_rng(123)
    J = (2 * rng.integers(0, 2, size=shape + (3,)) - 1).astype(np.float32)   # one disorder instance, many thermal histories

    def observables(res):
        hist = np.stack(res["overlap_histogram"]).astype(np.float64)
        coarse = np.stack([h[:512].reshape(16, 32).sum(axis=1) + np.eye(16)[15] * h[512] for h in hist])
        coarse /= coarse.sum(axis=1, keepdims=True)
        binder = 1 - res["overlap4"] / (3 * res["overlap2"] ** 2)
        return np.concatenate([res["energies"], res["overlap2"], binder]), coarse

    gpu_s, gpu_h, cpu_s, cpu_h = [], [], [], []
    for seed in range(n_seeds):
        g = pb.IsingSimulation(list(shape), J, temps, R, None, 9000 + seed, layout="int8")
        s, h = observables(g.sample(n_sweeps, "metropolis", pt_interval=1, warmup_ratio=0.25))
        gpu_s.append(s); gpu_h.append(h)
        c = oracle.Sim(shape, J, temps, n_replicas=R, seed=19000 + seed, rng_mode=oracle.RNG_XOSHIRO)
        s, h = observables(c.sample(n_sweeps, "metropolis", pt_interval=1, warmup_ratio=0.25))
        cpu_s.append(s); cpu_h.append(h)

    def z_of(a, b):
        a, b = np.asarray(a), np.asarray(b)
        sem = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
        return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(sem, 1e-12), sem

    z, _ = z_of(gpu_s, cpu_s)
    assert np.mean(np.abs(z) < 2.0) >= 0.9 and np.abs(z).max() < 3.5, z            # 12 numbers: 2 sigma, one outlier allowed
    zh, sem = z_of(gpu_h, cpu_h)
    live = sem > 1e-9                                                                # bins both runs never visit carry no information
    assert np.mean(np.abs(zh[live]) < 2.0) >= 0.9 and np.abs(zh[live]).max() < 4.0, zh
