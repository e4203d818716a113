#!/bin/bash
# r2 visit E: E-step / partition constants hoisted out of the tile CTAs; occupancy / cells-per-thread sweep (dev variants)
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_robust.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python scripts/est_probe.py default 2>&1 | tail -1
for v in variants/lib_*.so; do SQC_LIB=$PWD/$v timeout 300 python scripts/est_probe.py $v 2>&1 | tail -1; done
