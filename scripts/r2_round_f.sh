#!/bin/bash
# r2 visit F: full parity suite, the new bench line (parity_vs_oracle, fp64 co-limit, MLE / C4 legs), steady-state ncu of
# the E-step / partition kernels, per-leader k_mcd traces of a steady-state launch
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}.json"))
    r = d["roofline"]
    print("value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], "csteps", r["mcd_csteps_per_step"],
          {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
    print("fp64", r.get("fp64")); print("checks", d["checks"]); print("extra", json.dumps(d.get("extra"))[:1500]); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("failed", e)
PY
export SQC_BENCH_PROFILING=1
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra --rng counter"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_estep|k_partition' -s 60 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
timeout 300 python scripts/mcd_trace_all.py 1.0 30 1 > gpurun_out/${TAG}_trace_counter.log 2>&1; tail -3 gpurun_out/${TAG}_trace_counter.log
