#!/bin/bash
# 2 GPUs: robust-only timeline after the key-stream share / fence changes; parity
TAG=${1:-r2u}
mkdir -p gpurun_out
SQC_TIMELINE=gpurun_out/tl_${TAG}_w2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --no-e2e > gpurun_out/bench_${TAG}_w2.json 2> gpurun_out/bench_${TAG}_w2.err; tail -2 gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_w2.json"))
r = d["roofline"]
print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()}, d["checks"]["multi_gpu_parity"]["pass"])
PY
MGPU_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 scripts/mgpu_check.py 2>&1 | grep -E "MGPU CHECK|rng_state|z mismatches" | head
