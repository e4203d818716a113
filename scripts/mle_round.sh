#!/bin/bash
# single GPU: the MLE parity tests, then the MLE variant on the C3 cells with the in-tree library and build/ variants
TAG=${1:-mle}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mle.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -4 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
timeout 600 python scripts/mle_bench.py C3 20 2> gpurun_out/mle_bench_${TAG}.err | tee gpurun_out/mle_bench_${TAG}.json
for v in build/libsqc_*.so; do
  [ -f "$v" ] || continue
  SQC_LIB=$PWD/$v timeout 600 python scripts/mle_bench.py C3 20 2>> gpurun_out/mle_bench_${TAG}.err | tee -a gpurun_out/mle_bench_${TAG}.json
done
tail -3 gpurun_out/mle_bench_${TAG}.err
