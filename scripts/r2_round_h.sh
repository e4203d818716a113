#!/bin/bash
# r2 visit H: SMs reserved for the key-stream kernels beside k_mcd (R-compatible keys): sweep
TAG=${1:-r2k}
mkdir -p gpurun_out
for r in 0 24 40 50 60; do
SQC_MCD_RESERVE=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}_res$r.json 2> gpurun_out/bench_${TAG}_res$r.err
done
python - <<PY
import json
for r in (0, 24, 40, 50, 60):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_res%d.json" % r))
        rf = d["roofline"]
        print("reserve", r, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % rf["frac"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in rf["families"].items()})
    except Exception as e:
        print(r, "failed", e)
PY
SQC_TIMELINE=gpurun_out/tl_${TAG} timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
