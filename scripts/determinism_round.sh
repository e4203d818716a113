#!/bin/bash
# worker classes of k_mcd: default (classes stay apart: deterministic) against SQC_MCD_JOIN=1 (join when the other class
# has finished: timing decides when) — the bit-reproducibility test and the bench line's resident_equals_e2e_bitwise for both
TAG=${1:-r2det}
mkdir -p gpurun_out
for j in 0 1; do
  SQC_MCD_JOIN=$j timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k bit_reproducible > gpurun_out/pytest_${TAG}_join$j.log 2>&1; echo "join=$j: $(tail -1 gpurun_out/pytest_${TAG}_join$j.log)"
  SQC_MCD_JOIN=$j timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_${TAG}_join$j.json 2> gpurun_out/bench_${TAG}_join$j.err
done
python - <<PY
import json
for j in (0, 1):
    d = json.load(open("gpurun_out/bench_${TAG}_join%d.json" % j)); r = d["roofline"]
    print("join", j, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "bitwise", d["checks"].get("resident_equals_e2e_bitwise"),
          {k: round(v["avg_launch_ms"] * 1e3, 1) for k, v in r["families"].items()})
PY
