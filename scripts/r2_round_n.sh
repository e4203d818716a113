#!/bin/bash
# r2 visit N: MMD path, resident pipeline (arena), full suite
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}.json"))
    print("value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"]); print(json.dumps(d["extra"]["pipeline_C3"])[:1200])
except Exception as e:
    print("failed", e)
PY
timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from sampleqc_b200 import api
from oracle import oracle as O
rng = np.random.default_rng(0)
S, n_times, sub = 200, 20, 100
mats = [rng.normal(size=(int(n), 3)) + 0.05 * s for s, n in enumerate(rng.integers(300, 3000, size=S))]
api.calc_mmd_mat(mats[:10], n_times, sub, 3.0)
t = time.perf_counter(); m = api.calc_mmd_mat(mats, n_times, sub, 3.0); dt = time.perf_counter() - t
P = S * (S - 1) // 2
print(f"calc_mmd_mat: {S} samples, {P} pairs x {n_times} repetitions of {sub} x {sub}: {dt*1e3:.1f} ms ({P*n_times/dt:.3g} mmd_fn evaluations/s incl. host index draws)")
X, Y = mats[0][:100], mats[1][:100]
t = time.perf_counter(); [O.mmd_fn(X, Y, 3.0) for _ in range(200)]; dc = (time.perf_counter() - t) / 200
print(f"CPU restatement of mmd_fn(100 x 3, 100 x 3): {dc*1e6:.1f} us per call -> {1/dc:.3g} evaluations/s on one core")
PY
