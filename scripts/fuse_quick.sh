#!/bin/bash
# short single-GPU visit: core parity tests, then the bench line with a per-launch timeline
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robust.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
SQC_TIMELINE=gpurun_out/timeline_$TAG timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json")); r = d["roofline"]
print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "launches", d["gpu_launches"], {k: round(v["avg_launch_ms"] * 1e3, 1) for k, v in r["families"].items()})
PY
