"""Development check: CUDA fit vs oracle with per-output differences and timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from sampleqc_b200 import api
from sampleqc_b200.simulate import simulate_qcs, init_clusts

def run(N, J, K_true, K, D=3, seed=3, em=20, rng_mode=0, check=True):
    sim = simulate_qcs(N, J, K_true, D, seed=seed)
    iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    t0 = time.time()
    try:
        got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, 0.5, 50, 0, rng_mode=rng_mode)
    except Exception as e:
        print("GPU FAIL", N, J, K, e); got = None
    t1 = time.time()
    print(f"N={N} J={J} K={K} D={D} rng={rng_mode}: gpu e2e {t1-t0:.3f}s", "n_iter", got and got["n_iter"])
    if not check or got is None:
        return
    ref = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, 0.5, 50, 0, rng_mode=rng_mode)
    print(f"   oracle {time.time()-t1:.3f}s n_iter {ref['n_iter']}")
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2"):
        a, b = got[key], ref[key]
        print(f"   {key:8s} max rel diff {np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-300):.3e}")
    print("   z mismatches", int(np.sum(got["z"] != ref["z"])), "mcd_hit", got["mcd_iters_hit"], ref["mcd_iters_hit"])

if __name__ == "__main__":
    run(6000, 6, 3, 3)
    run(6000, 6, 3, 3, rng_mode=1)
    run(40000, 12, 4, 4)
    run(3000, 4, 3, 1)
    run(1_000_000, 100, 4, 4, seed=1002, rng_mode=0)
    run(1_000_000, 100, 4, 4, seed=1002, rng_mode=1)
