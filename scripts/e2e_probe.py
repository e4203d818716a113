import os, sys, time
os.environ["SQC_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
def pin(a):
    t = torch.empty(a.size, dtype={np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[a.dtype], pin_memory=True)
    v = t.numpy().reshape(a.shape, order="F" if a.ndim == 2 else "C"); v[...] = a; return v, t
xp, k1 = pin(x); gp, k2 = pin(g.astype(np.int32)); zp, k3 = pin(iz.astype(np.int32))
from sampleqc_b200 import _abi as A
for mode, (xx, gg, zz) in {"pinned": (xp, gp, zp)}.items():
    for i in range(4):
        t0 = time.perf_counter()
        fit = api.fit_sampleqc_robust_cpp(xx, zz, gg, D, J, K, N, 50, 0.5, 50, 0, rng_mode=0)
        print(mode, f"python-level total {1e3*(time.perf_counter()-t0):.1f} ms, n_iter {fit['n_iter']}", flush=True)
