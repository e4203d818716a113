"""small robust + MLE + outlier run for compute-sanitizer"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sampleqc_b200 import api
from sampleqc_b200.simulate import simulate_qcs, init_clusts
for (N, J, K, D) in [(30000, 9, 3, 3), (12000, 5, 2, 6)]:
    sim = simulate_qcs(N, J, K, D, seed=5)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=5)
    for rng in (0, 1):
        f = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 4, 0.5, 20, 0, rng_mode=rng)
    m = api.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, 2, 0)
    r = api.recentre(f)
    api.calc_outliers(sim.x, sim.groups, r["mu_0"], r["alpha_j"], r["beta_k"], r["sigma_k"], 0.01)
    print("ok", N, D, f["n_iter"], flush=True)
