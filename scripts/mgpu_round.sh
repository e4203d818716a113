#!/bin/bash
# multi-GPU visit: sharded parity check against the oracle, then the bench in both scaling modes
# usage: gpurun --gpus W -- bash scripts/mgpu_round.sh W <tag>
W=${1:-2}; TAG=${2:-r1}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511"
[ -z "$SKIP_CHECK" ] && timeout 600 $TR scripts/mgpu_check.py > gpurun_out/mgpu_check_${TAG}_w$W.log 2>&1; tail -4 gpurun_out/mgpu_check_${TAG}_w$W.log
[ -z "$SKIP_W1" ] && python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_w1.json 2> gpurun_out/bench_${TAG}_w1.err
for mode in weak strong; do
  [ "$mode" = strong ] && [ -n "$SKIP_STRONG" ] && continue
  for rng in ${RNGS:-r_compat counter}; do
    timeout 900 $TR bench.py --gpus $W --steps 3 --warmup 3 --no-cpu-baseline --scaling $mode --rng $rng > gpurun_out/bench_${TAG}_w${W}_${mode}_${rng}.json 2> gpurun_out/bench_${TAG}_w${W}_${mode}_${rng}.err
    tail -2 gpurun_out/bench_${TAG}_w${W}_${mode}_${rng}.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_w*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        fam = d["roofline"]["families"]
        print(f.split("/")[-1], "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "iters", d["em_iterations_per_step"], "e2e %.4g" % d["e2e"]["value"],
              {k: round(v["ms_per_step"] / d["em_iterations_per_step"], 3) for k, v in fam.items()})
    except Exception as e:
        print(f, "failed", e)
PY
