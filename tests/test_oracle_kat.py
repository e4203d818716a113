"""Pins the CPU oracle against every known answer available for this path:
R's RNG stream, qchisq, and structural properties of the fit.  CPU only."""
import numpy as np
import pytest
from scipy.stats import chi2

# set.seed(s); runif(3) — the familiar R outputs
R_RUNIF = {
    0: [0.8966972001362592, 0.2655086631421, 0.37212389963679016],
    1: [0.2655086631421, 0.37212389963679016, 0.5728533633518964],
    42: [0.9148060434963554, 0.9370754132978618, 0.2861395347863436],
    123: [0.28757752012461424, 0.7883051354438066, 0.40897692181169987],
}


@pytest.mark.parametrize("seed", sorted(R_RUNIF))
def test_r_runif_known_answers(oracle, seed):
    np.testing.assert_allclose(oracle.runif(seed, 3), R_RUNIF[seed], rtol=0, atol=1e-15)


def test_r_runif_printed_digits(oracle):
    # the 10-digit values quoted in SURVEY.md section 7
    assert round(oracle.runif(1, 1)[0], 10) == 0.2655086631
    assert round(oracle.runif(42, 1)[0], 10) == 0.9148060435
    assert round(oracle.runif(0, 1)[0], 10) == 0.8966972001


def test_randperm_keys_follow_runif(oracle):
    # Appendix A2: key = (int) Rf_runif(0, RAND_MAX)
    u = oracle.runif(7, 2000)
    k = oracle.rkeys(7, 0, 2000)
    np.testing.assert_array_equal(k, np.floor(2147483647.0 * u).astype(np.int64))
    # skipping draws is the same as slicing the stream (crosses two 624-word regenerations)
    np.testing.assert_array_equal(oracle.rkeys(7, 1500, 100), k[1500:1600])


@pytest.mark.parametrize("df", [1, 2, 3, 4, 5, 6, 7, 8])
def test_qchisq_against_scipy(oracle, df):
    ps = [1e-8, 1e-4, 0.01, 0.25, 0.4999, 0.5, 0.5001, 0.75, 0.95, 0.99, 0.999, 1 - 1e-9]
    got = np.array([oracle.qchisq(p, df) for p in ps])
    want = chi2.ppf(ps, df)
    np.testing.assert_allclose(got, want, rtol=5e-13)


def test_qchisq_quoted_values(oracle):
    assert oracle.qchisq(0.99, 3) == pytest.approx(11.344866730144373, rel=1e-14)
    assert oracle.qchisq(0.5, 3) == pytest.approx(2.3659738843753377, rel=1e-14)


def _small(seed=3, N=6000, J=6, K=3, D=3):
    from sampleqc_b200.simulate import simulate_qcs, init_clusts
    sim = simulate_qcs(N, J, K, D, seed=seed)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    return sim, iz, ig


def test_robust_fit_properties(oracle):
    sim, iz, _ = _small()
    N, D, J, K = sim.N, sim.D, sim.J, 3
    r = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 15, 0.5, 50, 0)
    assert list(r)[:12] == ["D", "J", "K", "N", "mu_0", "alpha_j", "beta_k", "sigma_k", "z", "p_jk", "like_1", "like_2"]
    assert r["mu_0"].shape == (D, 1) and r["alpha_j"].shape == (J, D) and r["beta_k"].shape == (K, D)
    assert r["sigma_k"].shape == (D, D, K) and r["z"].shape == (N, 1) and r["p_jk"].shape == (J, K)
    assert r["like_1"].shape == (15, 1)
    np.testing.assert_allclose(r["mu_0"].ravel(), sim.x.mean(axis=0), rtol=1e-13)
    np.testing.assert_allclose(r["p_jk"].sum(axis=1), 1.0, rtol=1e-13)
    assert np.all(np.diff(r["beta_k"][:, 0]) >= 0)              # ascending relabel :527-531
    zz = r["z"].ravel()
    assert zz.min() >= 1 and zz.max() <= K and np.all(zz == np.round(zz))
    for k in range(K):
        assert np.all(np.linalg.eigvalsh(r["sigma_k"][:, :, k]) > 0)
    n = r["n_iter"]
    assert np.all(r["like_1"].ravel()[n:] == 0) and np.all(r["like_2"].ravel()[n:] == 0)   # zero padding
    assert r["rng_draws"] == (1 + n) * N                        # Appendix A3
    # deterministic: same seed, same fit ("seeds are replicable", test-03_fit_sampleqc.R:81-88)
    r2 = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 15, 0.5, 50, 0)
    for key in ("alpha_j", "beta_k", "sigma_k", "p_jk", "z", "like_1", "like_2"):
        np.testing.assert_array_equal(r[key], r2[key])
    # a different seed gives a (slightly) different fit: the MCD start matters
    r3 = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 15, 0.5, 50, 1)
    assert not np.array_equal(r["beta_k"], r3["beta_k"])


def test_robust_z_agrees_with_mahalanobis(oracle):
    """the reference's intended-but-empty test 'z and mahalanobis distances agree' (test-03:90-91):
    with equal p_jk the label is the component with the largest log-density."""
    sim, iz, _ = _small(seed=11, N=4000, J=4, K=2)
    r = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, 4, 2, 4000, 10, 0.5, 50, 0)
    xc = sim.x - r["mu_0"].ravel() - r["alpha_j"][sim.groups]
    ll = np.empty((4000, 2))
    for k in range(2):
        S = r["sigma_k"][:, :, k]
        v = xc - r["beta_k"][k]
        q = np.einsum("ij,jk,ik->i", v, np.linalg.inv(S), v)
        ll[:, k] = np.log(r["p_jk"][sim.groups, k]) - 0.5 * np.log(np.linalg.det(S)) - 0.5 * q
    np.testing.assert_array_equal(np.argmax(ll, axis=1) + 1, r["z"].ravel().astype(int))


def test_k1_mcd_recovers_clean_gaussian(oracle):
    rng = np.random.default_rng(0)
    S = np.array([[1.0, 0.3, 0.1], [0.3, 0.5, -0.2], [0.1, -0.2, 2.0]])
    x = rng.multivariate_normal([1.0, -2.0, 0.5], S, size=40000)
    loc, scale, iters, q = oracle.fast_mcd(x, 20000, 50, 0)
    np.testing.assert_allclose(loc, [1.0, -2.0, 0.5], atol=0.03)
    np.testing.assert_allclose(scale, S, atol=0.06)            # consistency factor q/qchisq makes it unbiased
    assert 1 <= iters <= 50


def test_robust_error_classes(oracle):
    from sampleqc_b200._abi import SampleQCError, SQC_ERR_BAD_ARG, SQC_ERR_EMPTY_CLUSTER
    sim, iz, _ = _small(N=1000, J=2, K=2)
    with pytest.raises(SampleQCError) as e:
        oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, 2, 2, 1000, 5, 1.0, 50, 0)   # :471
    assert e.value.status == SQC_ERR_BAD_ARG
    with pytest.raises(SampleQCError) as e:
        oracle.fit_sampleqc_robust_cpp(sim.x, np.zeros(1000), sim.groups, 3, 2, 2, 1000, 5, 0.5, 50, 0)  # :307
    assert e.value.status == SQC_ERR_EMPTY_CLUSTER


def test_mle_fit_properties(oracle):
    sim, iz, ig = _small(seed=5)
    N, D, J, K = sim.N, sim.D, sim.J, 3
    m = oracle.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, 8, 0)
    assert list(m)[:14] == ["D", "J", "K", "N", "mu_0", "alpha_j", "beta_k", "sigma_k", "gamma_i", "z", "p_k", "p_jk", "like_1", "like_2"]
    np.testing.assert_allclose(m["gamma_i"].sum(axis=1), 1.0, rtol=1e-12)
    np.testing.assert_allclose(m["p_k"].sum(), 1.0, rtol=1e-12)
    np.testing.assert_allclose(m["p_jk"].sum(axis=1), 1.0, rtol=1e-12)
    assert np.all(np.diff(m["beta_k"][:, 0]) >= 0)
    l2 = m["like_2"].ravel()
    assert np.all(np.diff(l2) > -1e-6 * abs(l2[0]))              # EM monotone (up to the alpha step order)


def test_outlier_mask(oracle):
    sim, iz, _ = _small(seed=8, N=5000, J=5, K=2)
    r = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, 5, 2, 5000, 10, 0.5, 50, 0)
    # re-centring of R/fit_sampleqc.R:242-246
    am, bm = r["alpha_j"].mean(axis=0), r["beta_k"].mean(axis=0)
    mu0 = r["mu_0"].ravel() + am + bm
    aj, bk = r["alpha_j"] - am, r["beta_k"] - bm
    maha, out, cut = oracle.outlier_maha(sim.x, sim.groups, mu0, aj, bk, r["sigma_k"], 0.01)
    assert cut == pytest.approx(chi2.ppf(0.99, 3), rel=1e-13)
    xc = sim.x - mu0 - aj[sim.groups]
    for k in range(2):
        v = xc - bk[k]
        want = np.einsum("ij,jk,ik->i", v, np.linalg.inv(r["sigma_k"][:, :, k]), v)
        np.testing.assert_allclose(maha[:, k], want, rtol=1e-10)
    np.testing.assert_array_equal(out, maha.min(axis=1) > cut)
    assert 0.0 < out.mean() < 0.5


def test_init_clusts_oracle_pieces():
    """numpy restatement of .init_clusts' O(N) pieces (reference R/fit_sampleqc.R:286-290, :303-316): medians as R's
    median(), posterior against an independent implementation of the same mixture (sklearn)"""
    from sklearn.mixture import GaussianMixture
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    x = rng.normal(size=(3001, 3)) * [1.0, 2.0, 0.5] + [3.0, 2.5, -3.0]
    g = np.sort(rng.integers(0, 5, size=3001)).astype(np.int32)
    med = O.sample_medians(x, g, 6)
    assert np.isnan(med[5]).all()                                # no cells: median(numeric(0)) is NA
    for j in range(5):
        for d in range(3):
            v = np.sort(x[g == j, d]); n = v.size
            assert med[j, d] == (v[n // 2] if n % 2 else (v[n // 2 - 1] + v[n // 2]) / 2)
    med5 = med[:5]
    gm = GaussianMixture(3, covariance_type="full", random_state=0).fit(x - med5[g])
    sigma = np.asfortranarray(np.transpose(gm.covariances_, (1, 2, 0)))
    gamma, z = O.gmm_classify(x, g, med5, gm.weights_, gm.means_, sigma)
    np.testing.assert_allclose(gamma, gm.predict_proba(x - med5[g]), rtol=1e-9, atol=1e-13)
    np.testing.assert_array_equal(z, gm.predict(x - med5[g]))
    np.testing.assert_allclose(gamma.sum(axis=1), 1.0, rtol=1e-14)
