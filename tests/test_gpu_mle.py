"""GPU parity of the soft-EM variant (fit_sampleqc_mle_cpp) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _case(N, J, K_true, K, D=3, seed=3):
    from sampleqc_b200.simulate import simulate_qcs, init_clusts
    sim = simulate_qcs(N, J, K_true, D, seed=seed)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    return sim, iz, ig


@pytest.mark.parametrize("N,J,K_true,K,D,n_iter", [(6000, 6, 3, 3, 3, 10), (40000, 12, 4, 4, 3, 6), (9000, 5, 3, 2, 6, 5), (5000, 4, 3, 3, 3, 1)])
def test_mle_matches_oracle(oracle, N, J, K_true, K, D, n_iter):
    from sampleqc_b200 import api
    sim, _, ig = _case(N, J, K_true, K, D)
    ref = oracle.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, n_iter, 0)
    got = api.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, n_iter, 0)
    assert list(got)[:14] == list(ref)[:14] == ["D", "J", "K", "N", "mu_0", "alpha_j", "beta_k", "sigma_k", "gamma_i", "z", "p_k", "p_jk", "like_1", "like_2"]
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "like_1", "like_2"):
        scale = np.max(np.abs(ref[key]))
        np.testing.assert_allclose(got[key], ref[key], rtol=RTOL, atol=RTOL * scale, err_msg=key)
    np.testing.assert_allclose(got["gamma_i"], ref["gamma_i"], rtol=1e-7, atol=1e-10)
    # labels: identical wherever the two largest responsibilities are not within rounding of each other
    g = np.sort(ref["gamma_i"], axis=1)
    clear = (g[:, -1] - g[:, -2]) > 1e-9 if K > 1 else np.ones(N, bool)
    np.testing.assert_array_equal(got["z"].ravel()[clear], ref["z"].ravel()[clear])
    assert got["n_iter"] == n_iter


def test_mle_unsorted_groups(oracle):
    from sampleqc_b200 import api
    N, J, K = 6000, 6, 3
    sim, _, ig = _case(N, J, 3, K)
    perm = np.random.default_rng(1).permutation(N)
    x, g, g0 = np.asfortranarray(sim.x[perm]), sim.groups[perm], np.asfortranarray(ig[perm])
    ref = oracle.fit_sampleqc_mle_cpp(x, g0, g, 3, J, K, N, 4, 0)
    got = api.fit_sampleqc_mle_cpp(x, g0, g, 3, J, K, N, 4, 0)
    np.testing.assert_allclose(got["beta_k"], ref["beta_k"], rtol=RTOL)
    np.testing.assert_allclose(got["gamma_i"], ref["gamma_i"], rtol=1e-7, atol=1e-10)
    np.testing.assert_array_equal(got["z"], ref["z"])


def _mixture(N, J, K, D, seed, scale=1.0):
    """a plain Gaussian mixture with sample shifts, any D; responsibilities: a softened one-hot of the true labels"""
    rng = np.random.default_rng(seed)
    groups = np.sort(rng.integers(0, J, N)).astype(np.int32)
    z = rng.integers(0, K, N)
    x = rng.standard_normal((N, D)) * (0.5 + 0.2 * z[:, None]) + 3.0 * z[:, None] * ((np.arange(D) % 2) * 2 - 1) + 0.3 * rng.standard_normal((J, D))[groups]
    gam = np.full((N, K), 0.1 / K); gam[np.arange(N), z] += 0.9
    return np.asfortranarray(x * scale), groups, np.asfortranarray(gam)


def _cmp_mle(ref, got, K, N, rtol=RTOL):
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "like_1", "like_2"):
        scale = np.max(np.abs(ref[key]))
        if key == "beta_k" and K == 1:                            # one component: beta_k is the rounding noise around zero on both sides
            assert np.max(np.abs(got[key])) < 1e-12 and scale < 1e-12
            continue
        np.testing.assert_allclose(got[key], ref[key], rtol=rtol, atol=rtol * scale, err_msg=key)
    np.testing.assert_allclose(got["gamma_i"], ref["gamma_i"], rtol=1e-7, atol=1e-10)
    g = np.sort(ref["gamma_i"], axis=1)
    clear = (g[:, -1] - g[:, -2]) > 1e-9 if K > 1 else np.ones(N, bool)
    np.testing.assert_array_equal(got["z"].ravel()[clear], ref["z"].ravel()[clear])


@pytest.mark.parametrize("D,K", [(1, 2), (2, 3), (4, 2), (5, 3), (6, 6), (3, 1)])
def test_mle_every_dimension(oracle, D, K):
    """the moment sums leave the registers in groups of <= 16 values: 3, 6, 10, 15, 21 and 28 moments per component"""
    from sampleqc_b200 import api
    N, J = 7000, 5
    x, g, gam = _mixture(N, J, K, D, seed=40 + D)
    ref = oracle.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 4, 0)
    got = api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 4, 0)
    _cmp_mle(ref, got, K, N)


def test_mle_cells_far_from_every_component(oracle):
    """a likelihood under 1e-230 (still representable) leaves the table exponential for the reference's arithmetic and
    its true division; the fit must not notice"""
    from sampleqc_b200 import api
    N, J, K, D = 8000, 4, 2, 3
    x, g, gam = _mixture(N, J, K, D, seed=77)
    for i, far in zip((5, 4000, 7990), (13.0, 14.0, 15.0)):          # under the fitted parameters: likelihoods 2e-113, 2e-132 and 2e-273
        x[i] = x[i] + far * np.array([1.0, -1.0, 1.0])
    ref = oracle.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    got = api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    assert np.all(np.isfinite(ref["like_1"])) and np.min(ref["like_1"]) < 0
    _cmp_mle(ref, got, K, N)


def test_mle_weights_outside_the_fast_range(oracle):
    """x in units of 1e-9: the normaliser (2 pi)^(-D/2) det^(-1/2) ~ 1e27 is outside the range the table exponential's
    exponent arithmetic is safe for; every tile takes the reference's arithmetic"""
    from sampleqc_b200 import api
    N, J, K, D = 6000, 4, 2, 3
    x, g, gam = _mixture(N, J, K, D, seed=78, scale=1e-9)
    ref = oracle.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    got = api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 3, 0)
    _cmp_mle(ref, got, K, N)
