#!/bin/bash
# benches (both key modes, reference arm) + ncu launch list + full ncu captures, exported to CSV on the box (the .ncu-rep
# files are too large to travel back together); usage: bash scripts/gpu_evidence.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_rcompat.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --rng counter --no-cpu-baseline --no-extra > gpurun_out/bench_${TAG}_counter.json 2>> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
export SQC_BENCH_PROFILING=1
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 560 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_mcd|k_estep|k_partition|k_mt_jump|k_mt_gen' -s 150 -c 6 -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
for kn in k_mcd k_estep k_partition; do ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --kernel-name regex:$kn > gpurun_out/prof_${TAG}_src_$kn.csv 2>/dev/null; done
cat > /tmp/mle_probe.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
gam = np.full((N, K), 0.1 / K, order="F"); gam[np.arange(N), iz] += 0.9
s = api.Session(x, g, D, J, K, N, 6, variant=api.SQC_VARIANT_MLE, init_gamma=gam)
s.run(); s.run()
PY
python /tmp/mle_probe.py > gpurun_out/plain3_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_mle_pass' -s 8 -c 2 -o /tmp/prof_mle_$TAG python /tmp/mle_probe.py > gpurun_out/ncu_mle_$TAG.log 2>&1
ncu -i /tmp/prof_mle_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_mle_${TAG}_raw.csv 2>/dev/null
du -sh gpurun_out
python - <<PY
import json
for m in ("rcompat", "counter", "reference"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % m))
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "e2e", d.get("e2e", {}) and d["e2e"].get("value"), "roofline", d.get("roofline", {}).get("frac"))
        if m == "rcompat": print(json.dumps(d.get("extra", {}).get("pipeline_C3"))[:1500])
    except Exception as e:
        print(m, "failed", e)
PY
