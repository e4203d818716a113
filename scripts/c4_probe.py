"""per-family timing of C4 (5 M cells, D = 6, K = 6) on one GPU"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C4", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
s = api.Session(x, g, D, J, K, N, 8, init_z=iz, rng_mode=1)
s.profile(True); s.run(); s.run()
st = s.stats(); it = st["n_iter"]
print(sys.argv[1:], f"loop {st['loop_ms']:.2f} ms / {it} iters: estep {st['family_ms']['estep']/it*1e3:.0f} us  partition {st['family_ms']['partition']/(it+1)*1e3:.0f} us  mcd {st['family_ms']['mcd']/(it+1)*1e3:.0f} us  update {st['family_ms']['update']/it*1e3:.0f} us/iter  rounds {st['mcd_passes']}")
