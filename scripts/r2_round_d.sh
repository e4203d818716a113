#!/bin/bash
# r2 visit D: ncu full capture of the rewritten E-step / partition kernels; per-launch timeline of a 50-iteration fit
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 300 python scripts/est_probe.py $TAG 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_estep|k_partition' -s 6 -c 4 -o gpurun_out/prof_$TAG python scripts/est_probe.py > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
SQC_TIMELINE=gpurun_out/tl_$TAG timeout 300 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --rng counter > gpurun_out/bench_${TAG}_tl.json 2> gpurun_out/bench_${TAG}_tl.err
ls -la gpurun_out/prof_$TAG* gpurun_out/tl_$TAG*
