"""The resident pipeline (sqc_resident_*): .init_clusts -> fit -> re-centring -> outliers on one
upload of the cells, against the separate entry points and the CPU oracle (reference R/fit_sampleqc.R:195-262)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixture(x, groups, J, K, seed):
    """the caller's part of .init_clusts: a mixture fitted on a median-centred subsample (sklearn standing in for mclust)"""
    from sklearn.mixture import GaussianMixture
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(x.shape[0], size=min(x.shape[0], 10_000), replace=False))
    return idx, (lambda xs: GaussianMixture(K, covariance_type="full", random_state=seed).fit(xs))


@pytest.mark.parametrize("shuffle_codes", [False, True])
def test_resident_pipeline_matches_separate_steps_and_oracle(oracle, shuffle_codes):
    from sampleqc_b200 import api
    from sampleqc_b200.simulate import simulate_qcs
    N, J, K, D = 400_000, 40, 3, 3
    sim = simulate_qcs(N, J, 3, D, seed=21)
    g = sim.groups
    if shuffle_codes:                                             # runs in non-ascending code order, as factor() numbers them
        g = np.random.default_rng(5).permutation(J).astype(np.int32)[g]
    x = sim.x
    with api.Resident(x, g, J) as R:
        med = R.medians()
        ref_med = np.stack([np.median(x[g == j], axis=0) for j in range(J)])
        np.testing.assert_array_equal(med, ref_med)               # exact selection: bit-identical to R's / numpy's median
        idx, fit = _mixture(x, g, J, K, seed=4)
        xs = R.subsample(idx)
        np.testing.assert_array_equal(xs, x[idx] - ref_med[g[idx]])
        gm = fit(xs)
        sigma = np.ascontiguousarray(np.transpose(gm.covariances_, (1, 2, 0)))
        gamma, z = R.classify(gm.weights_, gm.means_, sigma, want_gamma=True, want_z=True)
        _, gamma2, z2 = api.gmm_classify(x, g, J, gm.weights_, gm.means_, sigma)
        np.testing.assert_array_equal(z, z2); np.testing.assert_array_equal(gamma, gamma2)
        assert np.unique(z).size == K
        got = R.fit_robust(K, 12, 0.5, 50, 0)                     # initial labels: the resident classification
        ref = api.recentre(oracle.fit_sampleqc_robust_cpp(x, z, g, D, J, K, N, 12, 0.5, 50, 0))
        assert got["n_iter"] == ref["n_iter"]
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
            np.testing.assert_allclose(np.asarray(got[key]).reshape(np.asarray(ref[key]).shape), ref[key], rtol=1e-8, atol=1e-8 * np.max(np.abs(ref[key])), err_msg=key)
        np.testing.assert_array_equal(got["z"], ref["z"])
        np.testing.assert_array_equal(got["rng_state"], ref["rng_state"])
        maha = np.zeros((N, K), order="F")
        m, flags, cut = R.outliers(0.01, out_maha=maha)
        m_ref, o_ref, c_ref = oracle.outlier_maha(x, g, ref["mu_0"], ref["alpha_j"], ref["beta_k"], ref["sigma_k"], 0.01)
        assert abs(cut - c_ref) < 1e-12 * c_ref
        np.testing.assert_allclose(m, m_ref, rtol=1e-7, atol=1e-9)
        near = np.abs(m_ref.min(axis=1) - c_ref) < 1e-9 * max(1.0, c_ref)
        np.testing.assert_array_equal(flags[~near], o_ref[~near])
        assert 0 < flags.sum() < N
        # the MLE variant from the resident responsibilities
        R.classify(gm.weights_, gm.means_, sigma, keep_gamma=True)
        gm_fit = R.fit_mle(K, 3, 0)
        ref_m = api.recentre(oracle.fit_sampleqc_mle_cpp(x, gamma, g, D, J, K, N, 3, 0))
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk"):
            np.testing.assert_allclose(np.asarray(gm_fit[key]).reshape(np.asarray(ref_m[key]).shape), ref_m[key], rtol=1e-8, atol=1e-8 * np.max(np.abs(ref_m[key])), err_msg=key)


def test_resident_rejects_interleaved_samples():
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    x = np.asfortranarray(np.random.default_rng(0).normal(size=(1000, 3)))
    g = np.tile(np.array([0, 1], dtype=np.int32), 500)
    with pytest.raises(SampleQCError) as e:
        api.Resident(x, g, 2)
    assert e.value.status == 1


def test_resident_refuses_steps_whose_inputs_are_not_on_the_handle():
    """outliers after a fit that was not re-centred, an MLE fit after a classification that kept no responsibilities,
    a robust fit with a K the resident labels were not made for: errors, not stale device buffers"""
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    from sampleqc_b200.simulate import simulate_qcs
    N, J, K, D = 40_000, 6, 2, 3
    sim = simulate_qcs(N, J, 2, D, seed=4)
    with api.Resident(sim.x, sim.groups, J) as R:
        idx, fit = _mixture(sim.x, sim.groups, J, K, 4)
        gm = fit(R.subsample(idx))
        with pytest.raises(SampleQCError):                        # no fit yet
            R.outliers(want_maha=False)
        sigma = np.ascontiguousarray(np.transpose(gm.covariances_, (1, 2, 0)))          # D x D x K, as mclust returns it
        R.classify(gm.weights_, gm.means_, sigma, keep_gamma=True)
        R.fit_mle(K, 3, 0)
        R.classify(gm.weights_, gm.means_, sigma, keep_gamma=False)
        with pytest.raises(SampleQCError) as e:
            R.fit_mle(K, 3, 0)                                    # the responsibilities of the first classification are not offered again
        assert e.value.status == 1
        with pytest.raises(SampleQCError) as e:
            R.fit_robust(K + 1, 3, 0.5, 20, 0)                    # labels are for K components
        assert e.value.status == 1
        R.fit_robust(K, 3, 0.5, 20, 0, recentre=False)
        with pytest.raises(SampleQCError) as e:
            R.outliers(want_maha=False)
        assert e.value.status == 1
        R.fit_robust(K, 3, 0.5, 20, 0)
        _, flags, cut = R.outliers(want_maha=False)
        assert flags.shape == (N,) and cut > 0
