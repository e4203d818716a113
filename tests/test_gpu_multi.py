"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): sharded robust + MLE fits over the in-kernel NVLink
exchange against the CPU oracle — parameters, labels, iteration counts and R's final .Random.seed (the key stream is
generated in per-rank shares).  Runs scripts/mgpu_check.py under torchrun, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fits_match_oracle_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MGPU_QUICK="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU CHECK OK" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("variant", ["robust", "mle"])
def test_one_call_n_gpus_matches_oracle(oracle, variant):
    """sqc_fit_args.n_gpus: ONE call, the library shards the cells over the GPUs of the process (host thread per GPU, peer
    access) — the path the Rcpp stub uses.  Sample runs in non-ascending code order, R-compatible keys; against the CPU
    oracle on the unsharded input."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from sampleqc_b200 import api
    from sampleqc_b200.simulate import init_clusts, simulate_qcs
    N, J, K, D = 300_000, 24, 3, 3
    sim = simulate_qcs(N, J, 4, D, seed=11)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=11)
    code = np.random.default_rng(1).permutation(J).astype(np.int32)
    g = code[sim.groups]
    if variant == "robust":
        ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, g, D, J, K, N, 8, 0.5, 50, 0, track=True)
        got = api.fit_sampleqc_robust_cpp(sim.x, iz, g, D, J, K, N, 8, 0.5, 50, 0, track=True, n_gpus=2)
        assert got["n_iter"] == ref["n_iter"]
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2", "track_alpha_j", "track_beta_k", "track_p_jk"):
            fin = np.isfinite(ref[key])                               # plain densities may underflow: log 0 = -inf on both sides
            np.testing.assert_array_equal(fin, np.isfinite(got[key]), err_msg=key)
            np.testing.assert_allclose(got[key][fin], ref[key][fin], rtol=1e-8, atol=1e-8 * np.max(np.abs(ref[key][fin])), err_msg=key)
        np.testing.assert_array_equal(got["z"], ref["z"])
        np.testing.assert_array_equal(got["rng_state"], ref["rng_state"])
        assert got["rng_draws"] == ref["rng_draws"]
    else:
        ref = oracle.fit_sampleqc_mle_cpp(sim.x, ig, g, D, J, K, N, 4, 0)
        got = api.fit_sampleqc_mle_cpp(sim.x, ig, g, D, J, K, N, 4, 0, n_gpus=2)
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "gamma_i", "like_1", "like_2"):
            np.testing.assert_allclose(got[key], ref[key], rtol=1e-7, atol=1e-7 * np.max(np.abs(ref[key])), err_msg=key)
        np.testing.assert_array_equal(got["z"], ref["z"])


def test_one_call_n_gpus_failure_is_agreed(oracle):
    """an error on one shard (a cluster with no cell anywhere: the reference's index error, :307) comes back as that
    error from the one call; nobody hangs"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    from sampleqc_b200.simulate import init_clusts, simulate_qcs
    N, J, K, D = 60_000, 8, 3, 3
    sim = simulate_qcs(N, J, 3, D, seed=2)
    iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=2)
    z_empty = np.where(iz == 2, 1, iz)
    with pytest.raises(SampleQCError) as e:
        api.fit_sampleqc_robust_cpp(sim.x, z_empty, sim.groups, D, J, K, N, 5, 0.5, 50, 0, n_gpus=2)
    assert e.value.status == 2
    bad = sim.groups.copy(); bad[N // 2 + 7] = J + 3               # bad input on the second shard: rejected before any GPU work
    with pytest.raises(SampleQCError) as e:
        api.fit_sampleqc_robust_cpp(sim.x, iz, bad, D, J, K, N, 5, 0.5, 50, 0, n_gpus=2)
    assert e.value.status == 1


def test_missing_peer_times_out_without_a_trap():
    """a rank whose peer never runs gives up with SQC_ERR_COMM after the wait budget; its CUDA context survives"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "scripts", "mgpu_timeout_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MGPU TIMEOUT CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_one_call_n_gpus_small_and_degenerate_inputs(oracle):
    """n_gpus > 1 on inputs the sharding has to adapt to: a single sample (one shard: the single-GPU path), fewer samples
    than GPUs, and a data set so small that the key stream is replicated instead of shared out"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from sampleqc_b200 import api
    from sampleqc_b200.simulate import init_clusts, simulate_qcs
    for N, J, K in ((20_000, 1, 2), (3_000, 2, 2), (9_000, 3, 3)):
        sim = simulate_qcs(N, J, K, 3, seed=31 + J)
        iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=31 + J)
        ref = oracle.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 6, 0.5, 50, 0)
        got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, 3, J, K, N, 6, 0.5, 50, 0, n_gpus=-1)      # SQC_ALL_GPUS
        assert got["n_iter"] == ref["n_iter"]
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
            np.testing.assert_allclose(got[key], ref[key], rtol=1e-8, atol=1e-8 * np.max(np.abs(ref[key])), err_msg=f"{key} N={N} J={J}")
        np.testing.assert_array_equal(got["z"], ref["z"])
        np.testing.assert_array_equal(got["rng_state"], ref["rng_state"])
