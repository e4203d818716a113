#!/bin/bash
# r2 visit J: full suite incl. the resident pipeline; reserve sweep around the default; bench line
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
for r in 16 24 32; do
SQC_MCD_RESERVE=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}_res$r.json 2> gpurun_out/bench_${TAG}_res$r.err
done
python - <<PY
import json
for r in (16, 24, 32):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_res%d.json" % r))
        rf = d["roofline"]
        print("reserve", r, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % rf["frac"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in rf["families"].items()})
    except Exception as e:
        print(r, "failed", e)
PY
