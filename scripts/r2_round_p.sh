#!/bin/bash
# r2 visit P (2 GPUs): multi-GPU parity after the IPC cache / retired key buffers, bench with one-wave key-stream shares
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_$TAG.log 2>&1; tail -15 gpurun_out/pytest_multi_$TAG.log | cut -c1-400
SQC_TIMELINE=gpurun_out/tl_${TAG}_w2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_w2.json 2> gpurun_out/bench_${TAG}_w2.err; tail -2 gpurun_out/bench_${TAG}_w2.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_w2.json"))
r = d["roofline"]
print("scaling", d["scaling"], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
o = d["extra"]["other_scaling_mode"]; print("weak", o["value"], o["ms_per_step"], o["e2e"]["value"], o["e2e"]["ms_per_step"], o["k_mcd_avg_launch_ms"])
print(d["checks"]["multi_gpu_parity"])
PY
