#!/bin/bash
# closing 2-GPU visit: the whole GPU suite (multi-GPU tests included), the reproducibility probe with sharded fits, the default bench line
TAG=${1:-r2z2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}_2gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}_2gpu.log
timeout 600 python scripts/determinism_probe.py C3 1.0 8 2>&1 | grep -v "^\[sqc\]" > gpurun_out/determinism_${TAG}.log; cat gpurun_out/determinism_${TAG}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_w2_strong_default.json 2> gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_w2_strong_default.json"))
print("2 GPUs", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["checks"].get("multi_gpu_parity"))
PY
