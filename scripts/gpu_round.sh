#!/bin/bash
# one GPU visit: parity tests, bench, launch list and a full ncu capture of the hot kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --rng counter > gpurun_out/bench_counter.json 2> gpurun_out/bench_counter.err; tail -2 gpurun_out/bench_counter.err; cat gpurun_out/bench_counter.json
export SQC_BENCH_PROFILING=1
CMD="python bench.py --steps 1 --warmup 1 --rng counter --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 470 -c 470 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mcd|k_estep|k_partition' -s 12 -c 6 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
