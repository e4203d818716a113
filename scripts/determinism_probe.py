"""How many distinct results do repeated fits of the same input give?  (bit-reproducibility probe, C3 by default)
usage: python scripts/determinism_probe.py [config] [scale] [repeats]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from sampleqc_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
x, g, iz, prm = workload(name, scale)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
KEYS = ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2", "z")


def digest(f):
    h = hashlib.sha1()
    for k in KEYS:
        h.update(np.ascontiguousarray(f[k]).tobytes())
    return h.hexdigest()[:10]


print("cluster sizes of the initial labels", np.bincount(iz, minlength=K).tolist(), "N", N, flush=True)
for rng_mode, rname in ((0, "r_compat"), (1, "counter")):
    for join in ("0", "1"):
        os.environ["SQC_MCD_JOIN"] = join
        fresh = [digest(api.fit_sampleqc_robust_cpp(x, iz, g, D, J, K, N, 50, 0.5, 50, 0, rng_mode=rng_mode)) for _ in range(reps)]
        with api.Session(x, g, D, J, K, N, 50, init_z=iz, mcd_alpha=0.5, mcd_iters=50, seed=0, rng_mode=rng_mode) as s:
            rerun = []
            for _ in range(reps):
                s.run(); rerun.append(digest(s.fetch()))
        print(f"{rname} join={join}: fresh fits {len(set(fresh))} distinct of {reps}; re-runs of one session {len(set(rerun))} distinct; "
              f"fresh == re-run: {set(fresh) == set(rerun)}  {sorted(set(fresh))[:4]} {sorted(set(rerun))[:4]}", flush=True)

# the MLE variant (fixed-order reductions only) and, with more than one GPU visible, the one-call sharded fits
os.environ["SQC_MCD_JOIN"] = "0"
gam = np.full((N, K), 0.1 / K, order="F"); gam[np.arange(N), iz] += 0.9
mk = ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "like_1", "like_2", "z")


def digest_mle(f):
    h = hashlib.sha1()
    for k in mk:
        if k in f:
            h.update(np.ascontiguousarray(f[k]).tobytes())
    return h.hexdigest()[:10]


mle = [digest_mle(api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 10, 0)) for _ in range(max(3, reps // 2))]
print(f"mle: {len(set(mle))} distinct of {len(mle)}", flush=True)
import torch
ng = torch.cuda.device_count()
if ng > 1:
    for join in ("0", "1"):
        os.environ["SQC_MCD_JOIN"] = join
        sh = [digest(api.fit_sampleqc_robust_cpp(x, iz, g, D, J, K, N, 50, 0.5, 50, 0, n_gpus=ng)) for _ in range(reps)]
        print(f"r_compat n_gpus={ng} join={join}: {len(set(sh))} distinct of {reps}", flush=True)
    os.environ["SQC_MCD_JOIN"] = "0"
    mle = [digest_mle(api.fit_sampleqc_mle_cpp(x, gam, g, D, J, K, N, 10, 0, n_gpus=ng)) for _ in range(max(3, reps // 2))]
    print(f"mle n_gpus={ng}: {len(set(mle))} distinct of {len(mle)}", flush=True)
