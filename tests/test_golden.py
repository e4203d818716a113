"""Committed golden fixtures (tests/golden/*.npz, made by the oracle with tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must reproduce them through the C ABI (GPU)."""
import glob
import os

import numpy as np
import pytest

ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLD = [p for p in ALL if not os.path.basename(p).startswith("init_")]          # fits
GOLD_INIT = [p for p in ALL if os.path.basename(p).startswith("init_")]         # .init_clusts pieces
KEYS_R = ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2")


def _load(path):
    d = np.load(path)
    return d, {k[4:]: d[k] for k in d.files if k.startswith("fit_")}


def _run(mod, d, name):
    if name.startswith("robust"):
        N, J, K, D, em, mi, seed, rng, track = [int(v) for v in d["params"]]
        return mod.fit_sampleqc_robust_cpp(np.asfortranarray(d["x"]), d["init_z"], d["groups"], D, J, K, N, em, float(d["mcd_alpha"]), mi, seed,
                                           track=bool(track), rng_mode=rng)
    N, J, K, D, it = [int(v) for v in d["params"]]
    return mod.fit_sampleqc_mle_cpp(np.asfortranarray(d["x"]), np.asfortranarray(d["init_gamma"]), d["groups"], D, J, K, N, it, 0)


def _check(got, want, name, rtol):
    keys = KEYS_R + (("gamma_i", "p_k") if name.startswith("mle") else ()) + (("track_beta_k", "track_sigma_k") if "track" in name else ())
    for k in keys:
        fin = np.isfinite(want[k])
        assert np.array_equal(fin, np.isfinite(got[k])), k
        scale = max(float(np.max(np.abs(want[k][fin]))), 1e-300) if fin.any() else 1.0
        np.testing.assert_allclose(got[k][fin], want[k][fin], rtol=rtol, atol=rtol * scale, err_msg=f"{name}:{k}")
    assert int(got["n_iter"]) == int(want["n_iter"])
    if name.startswith("robust"):
        np.testing.assert_array_equal(got["z"], want["z"])
        assert int(got["rng_draws"]) == int(want["rng_draws"])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(oracle, path):
    d, want = _load(path)
    name = os.path.basename(path)[:-4]
    _check(_run(oracle, d, name), want, name, rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_reproduces_golden(path):
    from sampleqc_b200 import api
    d, want = _load(path)
    name = os.path.basename(path)[:-4]
    got = _run(api, d, name)
    _check(got, want, name, rtol=1e-8)
    if name.startswith("robust") and int(d["params"][7]) == 0:
        np.testing.assert_array_equal(got["rng_state"], want["rng_state"])      # R's .Random.seed after the fit


@pytest.mark.parametrize("path", GOLD_INIT, ids=[os.path.basename(p)[:-4] for p in GOLD_INIT])
def test_oracle_reproduces_golden_init(oracle, path):
    d = np.load(path)
    N, J, K, D = [int(v) for v in d["params"]]
    med = oracle.sample_medians(d["x"], d["groups"], J)
    np.testing.assert_array_equal(med, d["medians"])
    gamma, z = oracle.gmm_classify(d["x"], d["groups"], med, d["pro"], d["mean"], d["sigma"])
    np.testing.assert_allclose(gamma, d["gamma"], rtol=1e-12, atol=1e-300)
    np.testing.assert_array_equal(z, d["z"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD_INIT, ids=[os.path.basename(p)[:-4] for p in GOLD_INIT])
def test_cuda_reproduces_golden_init(path):
    from sampleqc_b200 import api
    d = np.load(path)
    N, J, K, D = [int(v) for v in d["params"]]
    med, gamma, z = api.gmm_classify(np.asfortranarray(d["x"]), d["groups"], J, d["pro"], d["mean"], np.asfortranarray(d["sigma"]))
    np.testing.assert_array_equal(med, d["medians"])                            # medians: bit-exact
    np.testing.assert_allclose(gamma, d["gamma"], rtol=1e-10, atol=1e-14)
    top2 = np.sort(d["gamma"], axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-9
    np.testing.assert_array_equal(z[clear], d["z"][clear])


def test_fixture_set_is_complete():
    names = {os.path.basename(p)[:-4] for p in ALL}
    assert {"robust_toy_k2", "robust_k4_track", "robust_d6_k3", "robust_k1", "mle_k3", "mle_d6_k2", "init_k4", "init_d6_k3"} <= names
