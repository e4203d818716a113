#!/bin/bash
# one GPU visit: parity tests, bench (both key modes), launch list and a full ncu capture of the hot kernels
# usage: bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_rcompat.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --rng counter --no-cpu-baseline > gpurun_out/bench_${TAG}_counter.json 2>> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
export SQC_BENCH_PROFILING=1
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 560 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mcd|k_estep|k_partition|k_mt_jump|k_mt_gen' -s 20 -c 10 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
python - <<PY
import json
for m in ("rcompat", "counter", "reference"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % m))
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d.get("e2e", {}) and d["e2e"].get("value"), "roofline", d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(m, "failed", e)
PY
