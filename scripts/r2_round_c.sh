#!/bin/bash
# r2 visit C: parity of the rewritten E-step / partition kernels (dev build: D = 3, 6) + bench
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_robust.py tests/test_gpu_configs.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --rng counter > gpurun_out/bench_${TAG}_counter.json 2> gpurun_out/bench_${TAG}_counter.err
python - <<PY
import json
for m in ("", "_counter"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}%s.json" % m))
        r = d["roofline"]
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "iters", d["em_iterations_per_step"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], "csteps", r["mcd_csteps_per_step"],
              {k: round(v["ms_per_step"] / max(v["launches_per_step"], 1) * 1e3) for k, v in r["families"].items()}, d["checks"])
    except Exception as e:
        print(m, "failed", e)
PY
