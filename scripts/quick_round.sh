#!/bin/bash
# quick single-GPU visit (dev build D = 3, 6): core parity tests + both bench key modes
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_robust.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py -m gpu -q -x -k "not extreme_dims" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --rng counter > gpurun_out/bench_${TAG}_counter.json 2> gpurun_out/bench_${TAG}_counter.err
python - <<PY
import json
for m in ("", "_counter"):
    d = json.load(open("gpurun_out/bench_${TAG}%s.json" % m))
    r = d["roofline"]
    print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
PY
