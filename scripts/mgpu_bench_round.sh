#!/bin/bash
# N GPUs: sharded parity (process per GPU), bench --gpus N with the default flags (strong + weak extra, MLE, C5 at N = 8)
TAG=${1:-r2q}; NG=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
MGPU_QUICK=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 scripts/mgpu_check.py > gpurun_out/mgpu_check_${TAG}_w$NG.log 2>&1; grep -E "MGPU CHECK|robust N=|z mismatches|rng_state" gpurun_out/mgpu_check_${TAG}_w$NG.log | head -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_w$NG.json 2> gpurun_out/bench_${TAG}_w$NG.err; tail -3 gpurun_out/bench_${TAG}_w$NG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${TAG}_w$NG.json"))
    r = d["roofline"]
    print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
    print("checks", d["checks"].get("multi_gpu_parity")); 
    ex = d.get("extra", {})
    for k, v in ex.items(): print(k, json.dumps(v)[:700])
except Exception as e:
    print("failed", e)
PY
