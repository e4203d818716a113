#!/bin/bash
# A/B of the k_post_stats fusion (likelihood block + next iteration's alpha_j inside it) on one GPU:
# full GPU suite with the fusion on, core parity tests with it off, bench lines and a per-launch timeline for both
TAG=${1:-r2fuse}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
SQC_FUSE_ALPHA=0 timeout 600 python -m pytest tests/test_gpu_robust.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_gpu_${TAG}_off.log 2>&1; tail -2 gpurun_out/pytest_gpu_${TAG}_off.log
for f in 1 0; do
  SQC_FUSE_ALPHA=$f SQC_TIMELINE=gpurun_out/timeline_${TAG}_fuse$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}_fuse$f.json 2> gpurun_out/bench_${TAG}_fuse$f.err
done
python - <<PY
import json
for f in (1, 0):
    d = json.load(open("gpurun_out/bench_${TAG}_fuse%d.json" % f))
    r = d["roofline"]
    print("fuse", f, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "launches", d["gpu_launches"], {k: round(v["avg_launch_ms"] * 1e3, 1) for k, v in r["families"].items()})
PY
