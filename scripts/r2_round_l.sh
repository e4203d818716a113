#!/bin/bash
# r2 visit L: region-major candidate pass in the k_mcd leader post (dev build D = 3, 6): parity, probe, traces, bench
TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_robust.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py -m gpu -q -x -k "not extreme_dims" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --rng counter > gpurun_out/bench_${TAG}_counter.json 2> gpurun_out/bench_${TAG}_counter.err
python - <<PY
import json
for m in ("", "_counter"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}%s.json" % m))
        r = d["roofline"]
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
    except Exception as e:
        print(m, "failed", e)
PY
timeout 300 python scripts/mcd_trace_all.py 1.0 30 1 > gpurun_out/${TAG}_trace_counter.log 2>&1; grep -A11 "leader 3" gpurun_out/${TAG}_trace_counter.log | cut -c1-170
