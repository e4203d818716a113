"""Cross-check of the two independent CPU restatements of the reference:
the C++ oracle (oracle/sqc_oracle.cpp, loop for loop after the reference) against the numpy restatement
(oracle/np_restatement.py, written separately from the reference sources: LAPACK linear algebra, numpy's MT19937,
real sorts).  Bar: every output within 1e-9 relative, the same EM iteration count, identical labels.  CPU only."""
import numpy as np
import pytest

RTOL = 1e-9


def _case(N, J, K_true, K, D, seed):
    from sampleqc_b200.simulate import simulate_qcs, init_clusts
    sim = simulate_qcs(N, J, K_true, D, seed=seed)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    return sim, iz, ig


def test_r_stream_two_implementations(oracle):
    """R's set.seed + unif_rand through numpy's MT19937 == the oracle's own Mersenne-Twister, incl. known answers"""
    from oracle import np_restatement as R
    for seed in (0, 1, 42, 123, 2**31 + 5):
        np.testing.assert_array_equal(R.RUnif(seed).unif(3000), oracle.runif(seed, 3000))
        np.testing.assert_array_equal(R.RUnif(seed).randperm_keys(3000), oracle.rkeys(seed, 0, 3000))
    np.testing.assert_allclose(R.RUnif(42).unif(3), [0.9148060434963554, 0.9370754132978618, 0.2861395347863436], rtol=0, atol=1e-15)


@pytest.mark.parametrize("N,J,K_true,K,D,em,alpha,rng_mode,seed", [
    (3000, 4, 3, 3, 3, 15, 0.5, 0, 3),
    (20000, 8, 4, 2, 3, 12, 0.5, 0, 5),        # C1 shape: 4 true components fitted with K = 2
    (12000, 6, 4, 4, 3, 10, 0.6, 1, 7),
    (6000, 5, 3, 3, 6, 6, 0.5, 0, 9),          # D = 6
    (2500, 3, 3, 1, 3, 4, 0.75, 0, 11),
])
def test_robust_oracle_vs_numpy(oracle, N, J, K_true, K, D, em, alpha, rng_mode, seed):
    from oracle import np_restatement as R
    sim, iz, _ = _case(N, J, K_true, K, D, seed)
    a = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, alpha, 50, 1, rng_mode=rng_mode)
    b = R.fit_sampleqc_robust(sim.x, iz, sim.groups, D, J, K, N, em, alpha, 50, 1, rng_mode=rng_mode)
    assert a["n_iter"] == b["n_iter"]
    assert a["mcd_iters_hit"] == b["mcd_iters_hit"]
    np.testing.assert_array_equal(a["z"].ravel(), b["z"])
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
        ref = np.asarray(a[key]).reshape(b[key].shape)
        np.testing.assert_allclose(b[key], ref, rtol=RTOL, atol=RTOL * np.max(np.abs(ref)), err_msg=key)
    n = a["n_iter"]
    np.testing.assert_allclose(b["like_1"][:n], a["like_1"].ravel()[:n], rtol=RTOL)
    np.testing.assert_allclose(b["like_2"][:n], a["like_2"].ravel()[:n], rtol=RTOL)


@pytest.mark.parametrize("N,J,K,D,it,seed", [(5000, 5, 3, 3, 5, 13), (4000, 4, 2, 6, 3, 15)])
def test_mle_oracle_vs_numpy(oracle, N, J, K, D, it, seed):
    from oracle import np_restatement as R
    sim, _, ig = _case(N, J, K, K, D, seed)
    a = oracle.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, it, 0)
    b = R.fit_sampleqc_mle(sim.x, ig, sim.groups, D, J, K, N, it)
    np.testing.assert_array_equal(a["z"].ravel(), b["z"])
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "gamma_i", "like_1", "like_2"):
        ref = np.asarray(a[key]).reshape(b[key].shape)
        np.testing.assert_allclose(b[key], ref, rtol=RTOL, atol=RTOL * np.max(np.abs(ref)), err_msg=key)


def test_fast_mcd_oracle_vs_numpy(oracle):
    """the FastMCD alone (robust_cpp.cpp:55-108) on one cluster: location, scale, C-step count"""
    from oracle import np_restatement as R
    rng = np.random.default_rng(0)
    n, D = 5000, 3
    x = rng.standard_normal((n, D)) @ np.array([[1.0, 0.3, 0.0], [0.0, 0.7, -0.2], [0.0, 0.0, 0.4]])
    x[:400] += 6.0                                           # 8 % outliers
    h = int((n + D + 1) * 0.5)
    loc, scale, iters, q = oracle.fast_mcd(x, h, 50, 7)
    tr = []
    loc2, scale2, iters2 = R.fast_mcd(x, h, 50, R.RUnif(7).randperm_keys(n), tr)
    assert iters == iters2
    np.testing.assert_allclose(loc2, loc, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(scale2, scale, rtol=RTOL, atol=1e-12)
    assert tr[-1] == pytest.approx(q, rel=1e-10)
