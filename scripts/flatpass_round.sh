#!/bin/bash
# the flat candidate pass of the k_mcd leaders: core parity tests, the bench line (both key modes) and per-round leader traces
TAG=${1:-r2flat}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robust.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -2 gpurun_out/pytest_gpu_$TAG.log
for m in r_compat counter; do
  timeout 300 python bench.py --steps 5 --warmup 3 --rng $m --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}_$m.json 2> gpurun_out/bench_${TAG}_$m.err
done
timeout 200 python scripts/mcd_trace_all.py 1.0 30 0 2>&1 | grep -v "^\[sqc\]" > gpurun_out/mcd_trace_$TAG.log
python - <<PY
import json
for m in ("r_compat", "counter"):
    d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % m)); r = d["roofline"]
    print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % r["frac"], "rounds", r["mcd_rounds_per_step"], {k: round(v["avg_launch_ms"] * 1e3, 1) for k, v in r["families"].items()})
PY
grep -c . gpurun_out/mcd_trace_$TAG.log
