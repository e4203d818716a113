#!/bin/bash
# r2 visit B: full parity suite on the full build; bench A/B of the streaming E-step / partition kernels
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
SQC_STREAM=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_stream0.json 2> gpurun_out/bench_${TAG}_stream0.err
SQC_STREAM=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_stream1.json 2> gpurun_out/bench_${TAG}_stream1.err
python - <<PY
import json
for m in ("stream0", "stream1"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % m))
        r = d["roofline"]
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "iters", d["em_iterations_per_step"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], "csteps", r["mcd_csteps_per_step"],
              {k: round(v["ms_per_step"] / max(v["launches_per_step"], 1) * 1e3) for k, v in r["families"].items()})
    except Exception as e:
        print(m, "failed", e)
PY
