"""Edge cases of the CUDA path against the oracle: ragged / tiny samples, tile-boundary sample sizes, extreme D and K,
em_iters = 0, tied distances (duplicated cells), a single sample, unbalanced clusters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _cmp(ref, got):
    assert got["n_iter"] == ref["n_iter"]
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
        scale = max(np.max(np.abs(ref[key])), 1e-300)
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL, atol=RTOL * scale, err_msg=key)
    np.testing.assert_array_equal(got["z"], ref["z"])
    n = ref["n_iter"]
    for key in ("like_1", "like_2"):
        a, b = got[key].ravel()[:n], ref[key].ravel()[:n]
        fin = np.isfinite(b)
        assert np.array_equal(fin, np.isfinite(a))
        np.testing.assert_allclose(a[fin], b[fin], rtol=RTOL, err_msg=key)


def _gauss(rng, n_j, K, D, sep=3.0):
    """a plain mixture with given sample sizes (ragged allowed)"""
    J = len(n_j)
    groups = np.repeat(np.arange(J, dtype=np.int32), n_j)
    N = groups.size
    z = rng.integers(0, K, size=N).astype(np.int32)
    beta = rng.normal(size=(K, D)) * sep
    alpha = rng.normal(size=(J, D)) * 0.3
    x = np.asfortranarray(alpha[groups] + beta[z] + rng.normal(size=(N, D)) * (0.5 + 0.5 * rng.random(D)))
    return x, groups, z


@pytest.mark.parametrize("n_j", [[1024, 2048, 1, 2, 1023, 1025, 4097, 3], [7, 5, 9, 11, 6, 8, 300], [5000]])
def test_ragged_and_tile_boundary_samples(oracle, n_j):
    from sampleqc_b200 import api
    rng = np.random.default_rng(len(n_j))
    K, D = 2, 3
    x, g, z = _gauss(rng, n_j, K, D)
    N, J = x.shape[0], len(n_j)
    for mode in (0, 1):
        ref = oracle.fit_sampleqc_robust_cpp(x, z, g, D, J, K, N, 6, 0.5, 50, 0, rng_mode=mode)
        got = api.fit_sampleqc_robust_cpp(x, z, g, D, J, K, N, 6, 0.5, 50, 0, rng_mode=mode)
        _cmp(ref, got)


@pytest.mark.parametrize("D,K", [(1, 2), (2, 3), (4, 2), (5, 3), (7, 2), (8, 2), (3, 8), (3, 16)])
def test_extreme_dims(oracle, D, K):
    from sampleqc_b200 import api
    rng = np.random.default_rng(10 * D + K)
    n_j = [int(v) for v in rng.integers(400, 1500, size=6)]
    x, g, z = _gauss(rng, n_j, K, D, sep=6.0)
    N, J = x.shape[0], len(n_j)
    ref = oracle.fit_sampleqc_robust_cpp(x, z, g, D, J, K, N, 4, 0.5, 50, 0)
    got = api.fit_sampleqc_robust_cpp(x, z, g, D, J, K, N, 4, 0.5, 50, 0)
    _cmp(ref, got)
    gam = np.asfortranarray(np.eye(K)[z] * 0.9 + 0.1 / K)
    ref = oracle.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    got = api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    for key in ("alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "like_1", "like_2"):
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL, atol=RTOL * max(np.max(np.abs(ref[key])), 1e-300), err_msg=key)


def test_zero_iterations(oracle):
    from sampleqc_b200 import api
    rng = np.random.default_rng(3)
    x, g, z = _gauss(rng, [500, 700, 900], 3, 3)
    N = x.shape[0]
    ref = oracle.fit_sampleqc_robust_cpp(x, z, g, 3, 3, 3, N, 0, 0.5, 50, 0)
    got = api.fit_sampleqc_robust_cpp(x, z, g, 3, 3, 3, N, 0, 0.5, 50, 0)
    _cmp(ref, got)                                           # initial M-step only; z stays all-zero labels (:458)
    assert got["like_1"].size == 0 and got["rng_draws"] == N


def test_tied_distances_from_duplicated_cells(oracle):
    """exact ties at the h-subset boundary: the repo's tie rule (lower cell index first) must agree with the oracle's"""
    from sampleqc_b200 import api
    rng = np.random.default_rng(4)
    x, g, z = _gauss(rng, [600, 600], 2, 3)
    rep = 6
    x, g, z = np.asfortranarray(np.repeat(x, rep, axis=0)), np.repeat(g, rep), np.repeat(z, rep)   # every cell 6 times, contiguous by sample
    N = x.shape[0]
    for mode in (0, 1):
        ref = oracle.fit_sampleqc_robust_cpp(x, z, g, 3, 2, 2, N, 5, 0.5, 50, 0, rng_mode=mode)
        got = api.fit_sampleqc_robust_cpp(x, z, g, 3, 2, 2, N, 5, 0.5, 50, 0, rng_mode=mode)
        _cmp(ref, got)


def test_unbalanced_clusters_and_alpha_values(oracle):
    from sampleqc_b200 import api
    rng = np.random.default_rng(5)
    x, g, z = _gauss(rng, [3000, 2500, 4000, 1500], 3, 3, sep=5.0)
    z = np.where(rng.random(z.size) < 0.97, 0, z).astype(np.int32)       # one dominant cluster, two small ones
    z[:600] = np.arange(600) % 3                                          # the small clusters keep >= 200 cells (h <= n for mcd_alpha .95)
    N = x.shape[0]
    from sampleqc_b200._abi import SampleQCError
    outcomes = []
    for mcd_alpha in (0.5, 0.75, 0.95, 0.1):
        try:
            ref = oracle.fit_sampleqc_robust_cpp(x, z, g, 3, 4, 3, N, 5, mcd_alpha, 7, 0)
        except SampleQCError as e:
            # a component died out (or h > n) in the middle of the EM: the CUDA path must stop with the same failure class
            with pytest.raises(SampleQCError) as e2:
                api.fit_sampleqc_robust_cpp(x, z, g, 3, 4, 3, N, 5, mcd_alpha, 7, 0)
            assert e2.value.status == e.status
            outcomes.append(e.status)
            continue
        got = api.fit_sampleqc_robust_cpp(x, z, g, 3, 4, 3, N, 5, mcd_alpha, 7, 0)
        _cmp(ref, got)
        assert got["mcd_iters_hit"] == ref["mcd_iters_hit"]
        outcomes.append(0)
    assert 0 in outcomes                                     # at least one setting fits


def test_subset_size_errors_match(oracle):
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    rng = np.random.default_rng(6)
    x, g, z = _gauss(rng, [50, 60], 2, 3)
    N = x.shape[0]
    for fn in (oracle.fit_sampleqc_robust_cpp, api.fit_sampleqc_robust_cpp):
        with pytest.raises(SampleQCError) as e:              # h = int((n + D + 1) * 0) = 0: empty subset
            fn(x, z, g, 3, 2, 2, N, 3, 0.0, 10, 0)
        assert e.value.status == 3


def test_concurrent_single_gpu_fits_from_two_host_threads():
    """two host threads fitting different data on the same GPU at the same time get the results of the same calls made
    one after the other (per-session device state; the arena and the polynomial cache are the only shared objects)"""
    import threading
    from sampleqc_b200 import api
    from sampleqc_b200.simulate import init_clusts, simulate_qcs
    jobs = []
    for seed, (N, J, K) in enumerate(((150_000, 12, 3), (90_000, 7, 2))):
        sim = simulate_qcs(N, J, K, 3, seed=70 + seed)
        iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=70 + seed)
        jobs.append((sim, iz, ig, N, J, K))

    def run(job, variant):
        sim, iz, ig, N, J, K = job
        if variant == "robust":
            return api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 8, 0.5, 50, 0)
        return api.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, 3, J, K, N, 5, 0)

    plan = [(jobs[0], "robust"), (jobs[1], "robust"), (jobs[0], "mle")]
    alone = [run(j, v) for j, v in plan]
    for _ in range(3):
        res, errs = [None] * len(plan), []
        def work(i):
            try:
                res[i] = run(*plan[i])
            except Exception as e:                                # noqa: BLE001 - reported below
                errs.append(e)
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(plan))]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        for a, b in zip(alone, res):
            for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "z"):
                np.testing.assert_array_equal(a[key], b[key], err_msg=key)
