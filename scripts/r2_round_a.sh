#!/bin/bash
# r2 visit A: does the tree pass, and what do the streaming E-step / partition kernels buy (A/B by SQC_STREAM)
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
SQC_STREAM=0 timeout 300 python scripts/est_probe.py stream0 2>&1 | tail -2
SQC_STREAM=1 timeout 300 python scripts/est_probe.py stream1 2>&1 | tail -2
