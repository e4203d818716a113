"""MMD path: mmd_fn (reference src/mmd_fn.cpp:20-91) and the pair loop of .calc_mmd_mat
(reference R/calc_pairwise_mmds.R:198-280).  CPU: the oracle's loop-for-loop restatement against a vectorised numpy
formula.  GPU: the CUDA path against the oracle through the C ABI."""
import numpy as np
import pytest


def _np_mmd(X, Y, sigma):
    def k(A, B):
        d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
        return np.exp(-d2 / sigma)
    nx, ny = len(X), len(Y)
    kxx, kyy, kxy = k(X, X), k(Y, Y), k(X, Y)
    xx = np.triu(kxx, 1).sum() * (1.0 if nx == 1 else 2.0 / nx / (nx - 1))
    yy = np.triu(kyy, 1).sum() * (1.0 if ny == 1 else 2.0 / ny / (ny - 1))
    return xx + yy - 2 * kxy.sum() / ny / nx


def test_oracle_mmd_matches_vectorised_formula(oracle):
    rng = np.random.default_rng(0)
    for nx, ny, D, sigma in ((100, 100, 3, 3.0), (37, 250, 3, 3.0), (1, 40, 3, 1.5), (60, 1, 6, 6.0), (1, 1, 2, 2.0)):
        X, Y = rng.normal(size=(nx, D)), rng.normal(size=(ny, D)) + 0.3
        got, ref = oracle.mmd_fn(X, Y, sigma), _np_mmd(X, Y, sigma)
        assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref))
    from sampleqc_b200._abi import SampleQCError
    with pytest.raises(SampleQCError):
        oracle.mmd_fn(rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), 0.0)       # "sigma must be > 0", :28-30
    with pytest.raises(ValueError):
        oracle.mmd_fn(rng.normal(size=(5, 3)), rng.normal(size=(5, 2)), 1.0)       # :31-33


@pytest.mark.gpu
def test_mmd_fn_matches_oracle(oracle):
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    rng = np.random.default_rng(1)
    for nx, ny, D, sigma in ((100, 100, 3, 3.0), (129, 257, 3, 3.0), (1, 40, 3, 1.5), (60, 1, 6, 6.0), (1, 1, 2, 2.0), (1500, 900, 8, 8.0)):
        X, Y = rng.normal(size=(nx, D)), rng.normal(size=(ny, D)) * 1.2 + 0.3
        got, ref = api.mmd_fn(X, Y, sigma), oracle.mmd_fn(X, Y, sigma)
        assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref)), (nx, ny, D, got, ref)
    with pytest.raises(SampleQCError) as e:
        api.mmd_fn(rng.normal(size=(5, 3)), rng.normal(size=(5, 3)), -1.0)
    assert e.value.status == 1
    # identical samples: exactly the kernel's self-similarity structure, mmd ~ 0
    X = rng.normal(size=(80, 3))
    assert abs(api.mmd_fn(X, X, 3.0) - oracle.mmd_fn(X, X, 3.0)) < 1e-12


@pytest.mark.gpu
def test_mmd_pairs_match_the_reference_loop(oracle):
    """.calc_mmd_mat: every pair x n_times subsample repetitions in one device call == the R loop's sequence of
    abs(mmd_fn(m_i[sample(n_i, subsample), ], m_j[sample(n_j, subsample), ], sigma)) with the same draws"""
    from sampleqc_b200 import api
    rng = np.random.default_rng(2)
    sizes = [350, 90, 1200, 100, 101, 640]                           # some samples at / under the subsample size: used whole (:271, :276)
    mats = [rng.normal(size=(n, 3)) * (1 + 0.1 * s) + 0.2 * s for s, n in enumerate(sizes)]
    n_times, subsample, sigma = 5, 100, 3.0
    mat, allv, (pi, pj, idx_i, idx_j) = api.calc_mmd_mat(mats, n_times, subsample, sigma, seed=7, return_all=True)
    assert mat.shape == (6, 6) and np.allclose(mat, mat.T) and np.all(np.diag(mat) == 0)
    for p in range(pi.size):
        vals = []
        for t in range(n_times):
            mi = mats[pi[p]][idx_i[p, t]] if sizes[pi[p]] > subsample else mats[pi[p]]
            mj = mats[pj[p]][idx_j[p, t]] if sizes[pj[p]] > subsample else mats[pj[p]]
            vals.append(abs(oracle.mmd_fn(mi, mj, sigma)))                 # .dist_fn, :279
        np.testing.assert_allclose(allv[p], vals, rtol=1e-10, atol=1e-13)
        assert abs(mat[pi[p], pj[p]] - np.mean(vals)) <= 1e-12               # mean(mmd_vec), :231
