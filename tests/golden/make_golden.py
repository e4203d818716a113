"""Generates the committed golden fixtures tests/golden/*.npz with the CPU oracle (oracle/sqc_oracle.cpp).

The reference holds no numeric vectors for this path and cannot be built or run here (no R / Rcpp /
Armadillo), so these are ORACLE outputs ("parity unpinned"): they freeze the oracle's
behaviour so that a change of the restatement — or of the CUDA path, which must reproduce them too — shows.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402
from sampleqc_b200.simulate import init_clusts, simulate_qcs  # noqa: E402

CASES = {
    # name: (N, J, K_true, K, D, em_iters, mcd_alpha, mcd_iters, seed, rng_mode, data seed)
    "robust_toy_k2": (4000, 20, 4, 2, 3, 12, 0.5, 50, 0, 0, 101),      # toy-sized, K=2 robust (BASELINE config 1 in small)
    "robust_k4_track": (12000, 9, 4, 4, 3, 8, 0.5, 50, 1, 0, 102),     # seed = 1 is what track=TRUE sends
    "robust_d6_k3": (9000, 6, 3, 3, 6, 6, 0.6, 30, 0, 1, 103),         # CITE-seq style D = 6, counter keys
    "robust_k1": (3000, 5, 3, 1, 3, 5, 0.5, 50, 0, 0, 104),
}
MLE_CASES = {"mle_k3": (6000, 6, 3, 3, 3, 6, 105), "mle_d6_k2": (5000, 4, 3, 2, 6, 3, 106)}


def main():
    for name, (N, J, Kt, K, D, em, ma, mi, seed, rng, ds) in CASES.items():
        sim = simulate_qcs(N, J, Kt, D, seed=ds)
        iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=ds)
        track = "track" in name
        fit = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, ma, mi, seed, track=track, rng_mode=rng)
        fit = {k: v for k, v in fit.items() if isinstance(v, np.ndarray) or np.isscalar(v)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=sim.x.astype(np.float64), groups=sim.groups, init_z=iz,
                            params=np.array([N, J, K, D, em, mi, seed, rng, int(track)]), mcd_alpha=ma,
                            **{"fit_" + k: v for k, v in fit.items()})
        print(name, "n_iter", fit["n_iter"])
    for name, (N, J, Kt, K, D, it, ds) in MLE_CASES.items():
        sim = simulate_qcs(N, J, Kt, D, seed=ds)
        _, ig = init_clusts(sim.x, sim.groups, J, K, seed=ds)
        fit = O.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, it, 0)
        fit = {k: v for k, v in fit.items() if isinstance(v, np.ndarray) or np.isscalar(v)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=sim.x, groups=sim.groups, init_gamma=ig,
                            params=np.array([N, J, K, D, it]), **{"fit_" + k: v for k, v in fit.items()})
        print(name)
    # .init_clusts pieces (reference R/fit_sampleqc.R:286-316): per-sample medians, posterior and classification of
    # all cells under a mixture fitted on a subsample; the mixture's parameters are stored (no sklearn at test time)
    from sklearn.mixture import GaussianMixture
    for name, (N, J, Kt, K, D, ds) in {"init_k4": (6000, 11, 4, 4, 3, 107), "init_d6_k3": (4000, 5, 3, 3, 6, 108)}.items():
        sim = simulate_qcs(N, J, Kt, D, seed=ds)
        med = O.sample_medians(sim.x, sim.groups, J)
        idx = np.random.default_rng(ds).choice(N, 1500, replace=False)
        gm = GaussianMixture(K, covariance_type="full", random_state=ds).fit(sim.x[idx] - med[sim.groups[idx]])
        sigma = np.asfortranarray(np.transpose(gm.covariances_, (1, 2, 0)))
        gamma, z = O.gmm_classify(sim.x, sim.groups, med, gm.weights_, gm.means_, sigma)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=sim.x, groups=sim.groups, params=np.array([N, J, K, D]),
                            pro=gm.weights_, mean=gm.means_, sigma=sigma, medians=med, gamma=gamma, z=z)
        print(name)


if __name__ == "__main__":
    main()
