"""MLE (soft EM) timing on C3 and parity spot check at reduced scale"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sampleqc_b200 import api
from sampleqc_b200.simulate import make_config
sim, iz, ig, prm = make_config("C3")
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
s = api.Session(sim.x, sim.groups, D, J, K, N, 20, variant=api.SQC_VARIANT_MLE, init_gamma=ig)
s.profile(True); s.run(); s.run()
st = s.stats()
print(f"MLE C3 N={N}: n_iter {st['n_iter']} loop {st['loop_ms']:.1f} ms init {st['init_ms']:.1f} ms -> {N*st['n_iter']/st['loop_ms']*1e3:.3e} cell-iters/s; families", {k: round(v, 2) for k, v in st["family_ms"].items() if v}, {k: v for k, v in st["family_launches"].items() if v})
print("per pass: %.0f us" % (st["family_ms"]["mle_pass"] / st["family_launches"]["mle_pass"] * 1e3))
