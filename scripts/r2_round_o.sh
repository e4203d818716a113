#!/bin/bash
# r2 visit O (2 GPUs): per-launch timeline of a strong-scaling fit (where does k_partition's time go?), e2e with cached IPC mappings
TAG=${1:-r2s}
mkdir -p gpurun_out
SQC_TIMELINE=gpurun_out/tl_${TAG}_w2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 3 --no-extra > gpurun_out/bench_${TAG}_w2.json 2> gpurun_out/bench_${TAG}_w2.err; tail -2 gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_w2.json"))
r = d["roofline"]
print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
PY
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pipeline.py tests/test_gpu_init.py -m gpu -q 2>&1 | tail -3
