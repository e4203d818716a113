"""C4 (D=6, K=6) parity at reduced scale + C4 / C5 full-size sanity and timing on one GPU"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sampleqc_b200 import api
from sampleqc_b200.simulate import make_config
from oracle import oracle as O

sim, iz, ig, prm = make_config("C4", scale=0.04)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for rng in (0, 1):
    ref = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 10, 0.5, 50, 0, rng_mode=rng)
    got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 10, 0.5, 50, 0, rng_mode=rng)
    d = {k: float(np.max(np.abs(got[k] - ref[k])) / max(np.max(np.abs(ref[k])), 1e-300)) for k in ("alpha_j", "beta_k", "sigma_k", "p_jk")}
    print(f"C4 x0.04 N={N} rng={rng}: n_iter {got['n_iter']}/{ref['n_iter']} z mism {int(np.sum(got['z'] != ref['z']))} maxrel {max(d.values()):.2e}", flush=True)
for name in ("C4", "C5"):
    t0 = time.time()
    sim, iz, ig, prm = make_config(name)
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    print(name, "generated in", round(time.time() - t0, 1), "s", flush=True)
    for rng in (1, 0):
        s = api.Session(sim.x, sim.groups, D, J, K, N, 50, init_z=iz, rng_mode=rng)
        s.profile(True); s.run(); s.run()
        st = s.stats(); fit = s.fetch()
        ok = np.all(np.isfinite(fit["beta_k"])) and np.allclose(fit["p_jk"].sum(axis=1), 1.0) and fit["z"].min() >= 1 and fit["z"].max() <= K
        print(f"{name} N={N} D={D} K={K} rng={rng}: n_iter {st['n_iter']} loop {st['loop_ms']:.1f} ms -> {N*st['n_iter']/st['loop_ms']*1e3:.3e} cell-iters/s sane={bool(ok)} "
              f"fam={ {k: round(v, 1) for k, v in st['family_ms'].items() if v} } rounds={st['mcd_passes']}", flush=True)
        s.close()
