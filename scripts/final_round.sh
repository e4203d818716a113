#!/bin/bash
# the round's closing single-GPU visit: full GPU suite, smoke, then the evidence (benches, ncu)
TAG=${1:-r2final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_evidence.sh $TAG
