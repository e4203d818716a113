#!/bin/bash
# r2 visit I (2 GPUs): process-per-GPU parity (mgpu_check), one-call n_gpus parity, bench --gpus 2 (strong default + weak extra)
TAG=${1:-r2l}
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$TAG.log 2>&1; tail -15 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_w2.json 2> gpurun_out/bench_${TAG}_w2.err; tail -3 gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_w2.json"))
    r = d["roofline"]
    print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d["e2e"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
    print("checks", d["checks"]); print("extra", json.dumps(d.get("extra"))[:1800])
except Exception as e:
    print("failed", e)
PY
# one-call path timing at C3: 1 vs 2 GPUs through sqc_fit_robust (host buffers)
timeout 600 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for n in (1, 2):
    for rep in range(3):
        t0 = time.perf_counter()
        f = api.fit_sampleqc_robust_cpp(x, iz, g, D, J, K, N, 50, 0.5, 50, 0, n_gpus=n)
        dt = time.perf_counter() - t0
    print(f"one call n_gpus={n}: {dt*1e3:.1f} ms, {f['n_iter']} iterations, {N*f['n_iter']/dt:.4g} cell-iters/s", flush=True)
    if n == 1: ref = f
    else: print("  max rel diff vs 1 GPU:", max(float(np.max(np.abs(f[k]-ref[k]))/np.max(np.abs(ref[k]))) for k in ("alpha_j","beta_k","sigma_k","p_jk")), "z equal", bool(np.array_equal(f["z"], ref["z"])))
PY
