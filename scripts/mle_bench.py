"""MLE variant (fit_sampleqc_mle_cpp) on the C3 cells: ms per fit and per kernel family, for the library SQC_LIB points to.

    SQC_LIB=/path/to/lib.so python scripts/mle_bench.py [workload] [n_iter]
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                      # noqa: E402  (device memory / synchronize only)
from sampleqc_b200 import api                                     # noqa: E402
from sampleqc_b200.simulate import make_config                    # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sim, iz, _, prm = make_config(wl, scale=float(os.environ.get("SCALE", "1")))
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
gam = np.full((N, K), 0.1 / K, order="F"); gam[np.arange(N), iz] += 0.9
sess = api.Session(sim.x, sim.groups.astype(np.int32), D, J, K, N, n_iter, variant=api.SQC_VARIANT_MLE, init_gamma=gam, seed=0, device=0)
sess.profile(True)
for _ in range(2):
    sess.run()
torch.cuda.synchronize()
steps, ms, fam, cnt = 3, 0.0, {}, {}
for _ in range(steps):
    sess.run()
    st = sess.stats()
    ms += st["loop_ms"] + st["init_ms"]
    for k, v in st["family_ms"].items():
        fam[k] = fam.get(k, 0.0) + v; cnt[k] = cnt.get(k, 0) + st["family_launches"][k]
sess.close()
tl = os.environ.get("SQC_TIMELINE")
if tl and os.path.exists(tl + ".rank0.csv"):                     # per-tag launch durations and the gaps behind them (last run)
    import csv
    rows = list(csv.DictReader(open(tl + ".rank0.csv")))
    agg = {}
    for a, b in zip(rows, rows[1:] + [None]):
        d = agg.setdefault(a["tag"] or a["family"], [0, 0.0, 0.0])
        d[0] += 1; d[1] += float(a["t1_ms"]) - float(a["t0_ms"])
        if b is not None:
            d[2] += float(b["t0_ms"]) - float(a["t1_ms"])
    print(json.dumps({"timeline_us": {k: {"n": v[0], "avg": round(1e3 * v[1] / v[0], 1), "gap_after": round(1e3 * v[2] / v[0], 1)} for k, v in agg.items()}}))
print(json.dumps({"lib": os.environ.get("SQC_LIB", "in-tree"), "workload": wl, "n_iter": n_iter, "ms_per_fit": ms / steps, "cell_iters_per_s": N * n_iter / (ms / steps / 1e3),
                  "family_ms_per_fit": {k: round(v / steps, 4) for k, v in fam.items() if v}, "avg_launch_us": {k: round(1e3 * fam[k] / cnt[k], 1) for k in fam if cnt[k]}}))
