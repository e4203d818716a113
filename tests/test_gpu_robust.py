"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star): parameters within 1e-8 relative, the same number of EM
iterations, identical labels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-8


def _case(N, J, K_true, K, D=3, seed=3):
    from sampleqc_b200.simulate import simulate_qcs, init_clusts
    sim = simulate_qcs(N, J, K_true, D, seed=seed)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    return sim, iz, ig


def _compare_robust(ref, got, N, rtol=RTOL):
    assert got["n_iter"] == ref["n_iter"]
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
        scale = np.max(np.abs(ref[key]))
        np.testing.assert_allclose(got[key], ref[key], rtol=rtol, atol=rtol * scale, err_msg=key)
    n = ref["n_iter"]
    np.testing.assert_allclose(got["like_1"][:n], ref["like_1"][:n], rtol=rtol, err_msg="like_1")
    np.testing.assert_allclose(got["like_2"][:n], ref["like_2"][:n], rtol=rtol, err_msg="like_2")
    assert np.all(got["like_1"][n:] == 0) and np.all(got["like_2"][n:] == 0)
    np.testing.assert_array_equal(got["z"], ref["z"])
    assert got["rng_draws"] == ref["rng_draws"] == (1 + n) * N
    assert got["mcd_iters_hit"] == ref["mcd_iters_hit"]


@pytest.mark.parametrize("rng_mode", [0, 1])
@pytest.mark.parametrize("N,J,K_true,K", [(6000, 6, 3, 3), (40000, 12, 4, 4), (20000, 5, 4, 2), (3000, 4, 3, 1)])
def test_robust_matches_oracle(oracle, N, J, K_true, K, rng_mode):
    from sampleqc_b200 import api
    sim, iz, _ = _case(N, J, K_true, K)
    ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 20, 0.5, 50, 0, rng_mode=rng_mode)
    got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 20, 0.5, 50, 0, rng_mode=rng_mode)
    _compare_robust(ref, got, N)
    if rng_mode == 0:
        np.testing.assert_array_equal(got["rng_state"], ref["rng_state"])


def test_robust_track_and_seed(oracle):
    from sampleqc_b200 import api
    N, J, K = 8000, 5, 3
    sim, iz, _ = _case(N, J, 3, K, seed=11)
    ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 8, 0.6, 30, 1, track=True)
    got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 8, 0.6, 30, 1, track=True)
    _compare_robust(ref, got, N)
    n = ref["n_iter"]
    for key in ("track_alpha_j", "track_beta_k", "track_sigma_k", "track_p_jk"):
        np.testing.assert_allclose(got[key][..., :n], ref[key][..., :n], rtol=RTOL, atol=1e-12, err_msg=key)


def test_robust_deterministic():
    from sampleqc_b200 import api
    N, J, K = 30000, 7, 3
    sim, iz, _ = _case(N, J, 3, K, seed=5)
    a = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 10, 0.5, 50, 0)
    b = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 10, 0.5, 50, 0)
    for key in ("alpha_j", "beta_k", "sigma_k", "p_jk", "z", "like_1", "like_2"):
        np.testing.assert_array_equal(a[key], b[key])       # "seeds are replicable", test-03_fit_sampleqc.R:81-88


def test_robust_unsorted_groups(oracle):
    """cells of a sample interleaved: sorted by sample once; with counter keys the stream position follows the
    sample-sorted storage order, so compare with the oracle on the stably sorted data and map the labels back"""
    from sampleqc_b200 import api
    N, J, K = 6000, 6, 3
    sim, iz, _ = _case(N, J, 3, K)
    perm = np.random.default_rng(0).permutation(N)
    x, g, z0 = np.asfortranarray(sim.x[perm]), sim.groups[perm], iz[perm]
    order = np.argsort(g, kind="stable")
    for mode in (0, 1):
        ref = oracle.fit_sampleqc_robust_cpp(np.asfortranarray(x[order]), z0[order], g[order], 3, J, K, N, 10, 0.5, 50, 0, rng_mode=mode)
        got = api.fit_sampleqc_robust_cpp(x, z0, g, 3, J, K, N, 10, 0.5, 50, 0, rng_mode=mode)
        np.testing.assert_allclose(got["beta_k"], ref["beta_k"], rtol=RTOL)
        np.testing.assert_array_equal(got["z"].ravel()[order], ref["z"].ravel())


@pytest.mark.parametrize("variant", ["robust", "mle"])
def test_sample_runs_in_non_ascending_code_order(oracle, variant):
    """what fit_sampleqc() sends (R/fit_sampleqc.R:205-216): x is an rbind over samples, but groups =
    as.integer(factor(sample_ids)) numbers the samples alphabetically — contiguous runs whose codes are not ascending.
    The reference pairs the m-th randperm key with the m-th cell of a cluster in the CALLER's cell order; compared with
    the oracle on the unsorted input itself, R-compatible keys, plus R's RNG state after the fit."""
    from sampleqc_b200 import api
    N, J, K = 30_000, 7, 3
    sim, iz, ig = _case(N, J, 3, K, seed=9)
    code = np.array([4, 0, 6, 2, 5, 1, 3], dtype=np.int32)        # run r carries sample code code[r]
    g = code[sim.groups]
    assert np.any(np.diff(g) < 0)
    if variant == "robust":
        ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, g, 3, J, K, N, 10, 0.5, 50, 0, track=True)
        got = api.fit_sampleqc_robust_cpp(sim.x, iz, g, 3, J, K, N, 10, 0.5, 50, 0, track=True)
        assert got["n_iter"] == ref["n_iter"]
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "track_alpha_j", "track_p_jk"):
            np.testing.assert_allclose(got[key], ref[key], rtol=RTOL, atol=RTOL * np.max(np.abs(ref[key])), err_msg=key)
        np.testing.assert_array_equal(got["z"], ref["z"])
        np.testing.assert_array_equal(got["rng_state"], ref["rng_state"])
    else:
        ref = oracle.fit_sampleqc_mle_cpp(sim.x, ig, g, 3, J, K, N, 4, 0)
        got = api.fit_sampleqc_mle_cpp(sim.x, ig, g, 3, J, K, N, 4, 0)
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "gamma_i"):
            np.testing.assert_allclose(got[key], ref[key], rtol=RTOL, atol=RTOL * np.max(np.abs(ref[key])), err_msg=key)
        np.testing.assert_array_equal(got["z"], ref["z"])


def test_error_classes(oracle):
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    N, J, K = 3000, 3, 3
    sim, iz, _ = _case(N, J, 3, K)
    with pytest.raises(SampleQCError) as e:                  # robust_cpp.cpp:471-472
        api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 5, 1.0, 50, 0)
    assert e.value.status == 1
    z_empty = np.where(iz == 2, 1, iz)                       # empty cluster -> x_k.rows(0,-1), :307
    with pytest.raises(SampleQCError) as e:
        api.fit_sampleqc_robust_cpp(sim.x, z_empty, sim.groups, 3, J, K, N, 5, 0.5, 50, 0)
    assert e.value.status == 2
    with pytest.raises(SampleQCError) as e2:
        oracle.fit_sampleqc_robust_cpp(sim.x, z_empty, sim.groups, 3, J, K, N, 5, 0.5, 50, 0)
    assert e2.value.status == e.value.status
    xn = sim.x.copy(); xn[17, 1] = np.nan                    # NaN reaches sort_index, :50
    with pytest.raises(SampleQCError):
        api.fit_sampleqc_robust_cpp(xn, iz, sim.groups, 3, J, K, N, 5, 0.5, 50, 0)


def test_device_special_functions(oracle):
    from scipy.stats import chi2
    from sampleqc_b200 import api
    ps = np.array([1e-8, 1e-4, 0.01, 0.25, 0.4999, 0.5, 0.5001, 0.75, 0.95, 0.99, 0.999, 1 - 1e-9])
    for df in range(1, 9):
        np.testing.assert_allclose(api.device_qchisq(ps, df), chi2.ppf(ps, df), rtol=5e-13)
    np.testing.assert_array_equal(api.device_rkeys(7, 0, 5000), oracle.rkeys(7, 0, 5000))
    np.testing.assert_array_equal(api.device_rkeys(0, 1500, 3000), oracle.rkeys(0, 1500, 3000))
    # several 68 640-word blocks: exercises the GF(2) jump-ahead table and the state hand-over between calls
    np.testing.assert_array_equal(api.device_rkeys(1, 100_000, 400_000), oracle.rkeys(1, 100_000, 400_000))
    # long blocks (the production setting: ~16 blocks of several hundred chunks per call)
    np.testing.assert_array_equal(api.device_rkeys(2, 77_777, 3_000_000), oracle.rkeys(2, 77_777, 3_000_000))
    np.testing.assert_array_equal(api.device_rkeys(5, 623, 1), oracle.rkeys(5, 623, 1))
    np.testing.assert_array_equal(api.device_rkeys(5, 624, 625), oracle.rkeys(5, 624, 625))


@pytest.mark.parametrize("n_total,world,calls,seed", [(1_000_003, 3, 3, 11), (5_000_000, 8, 2, 0), (40_000, 2, 4, 3), (624 * 16, 2, 3, 5)])
def test_distributed_key_stream_matches_r(oracle, n_total, world, calls, seed):
    """multi-GPU fits generate R's key stream in per-rank shares (host-derived GF(2) jump polynomials move the
    Mersenne-Twister state from call to call and to a rank's offset): the shares of all ranks, concatenated, are the
    stream set.seed(seed) gives, call after call, and every rank ends with the same state"""
    from sampleqc_b200 import api
    want = oracle.rkeys(seed, 0, calls * n_total).reshape(calls, n_total)
    states = []
    for r in range(world):
        got, q, st = api.device_rkeys_share(seed, n_total, world, r, calls)
        lo, hi = r * q, min((r + 1) * q, n_total)
        for c in range(calls):
            np.testing.assert_array_equal(got[c, :max(hi - lo, 0)], want[c, lo:hi])
        states.append(st)
    for st in states[1:]:
        np.testing.assert_array_equal(st, states[0])
    # the state after the last call continues the stream: one more call from it through the single-stream generator
    assert states[0][624] >= 1 and states[0][624] <= 624


def test_outlier_mask_matches_oracle(oracle):
    from sampleqc_b200 import api
    N, J, K = 20000, 6, 3
    sim, iz, _ = _case(N, J, 3, K)
    fit = api.recentre(api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 10, 0.5, 50, 0))
    m_ref, o_ref, c_ref = oracle.outlier_maha(sim.x, sim.groups, fit["mu_0"], fit["alpha_j"], fit["beta_k"], fit["sigma_k"], 0.01)
    m, o, c = api.calc_outliers(sim.x, sim.groups, fit["mu_0"], fit["alpha_j"], fit["beta_k"], fit["sigma_k"], 0.01)
    assert c == pytest.approx(c_ref, rel=1e-13)
    np.testing.assert_allclose(m, m_ref, rtol=1e-10)
    near = np.abs(m_ref.min(axis=1) - c_ref) < 1e-9          # flags identical apart from cells within 1e-9 of the cut
    np.testing.assert_array_equal(o[~near], o_ref[~near])
    assert o.sum() > 0


def test_device_exp_matches_libm():
    """the kernels' exp(x), x <= 0: <= 2.5 ulp over the normal range, correct gradual underflow, NaN / -inf handling"""
    from sampleqc_b200 import api
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(200000) * 700.0, -rng.random(20000) * 1e-3, -np.exp(rng.normal(size=20000) * 3), [0.0, -0.0, -708.0, -708.4, -1e-300]])
    got, want = api.device_expneg(x), np.exp(x)
    ulp = np.abs(got - want) / np.spacing(want)
    assert ulp.max() <= 2.5, ulp.max()
    xs = -np.linspace(708.0, 760.0, 5000)                    # subnormal results and underflow to zero
    got, want = api.device_expneg(xs), np.exp(xs)
    np.testing.assert_allclose(got, want, rtol=0, atol=3 * 4.95e-324 + 6e-16 * want.max())
    assert np.all(np.abs(got - want) <= 3 * np.maximum(np.spacing(want), 4.95e-324))
    sp = api.device_expneg(np.array([-np.inf, -1e6, np.nan]))
    assert sp[0] == 0.0 and sp[1] == 0.0 and np.isnan(sp[2])


@pytest.mark.parametrize("rng_mode", [0, 1])
def test_robust_large_clusters(oracle, rng_mode):
    """clusters above 65 536 cells: the predicted-band (direct COLLECT) rounds of the key start and of the settled
    C-steps, several chunks per worker, the block-parallel key stream"""
    from sampleqc_b200 import api
    N, J, K = 400_000, 20, 2
    sim, iz, _ = _case(N, J, 4, K, seed=21)
    ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 4, 0.5, 50, 0, rng_mode=rng_mode)
    got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 4, 0.5, 50, 0, rng_mode=rng_mode)
    _compare_robust(ref, got, N)
