#!/bin/bash
# 2 GPUs: the whole GPU suite on a 2-GPU box (multi tests run), timeout check, default 2-GPU bench
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}_2gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu_${TAG}_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_w2.json 2> gpurun_out/bench_${TAG}_w2.err; tail -2 gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_w2.json"))
r = d["roofline"]
print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
o = d["extra"]["other_scaling_mode"]; print("weak", o["value"], o["ms_per_step"], o["e2e"]["value"], o["e2e"]["ms_per_step"], o["k_mcd_avg_launch_ms"])
print(d["checks"]["multi_gpu_parity"]["pass"])
PY
