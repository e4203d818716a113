#!/bin/bash
# r2 visit K: leader-only incremental rounds in k_mcd (dev build D = 3, 6): parity, A/B, traces
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_robust.py tests/test_golden.py tests/test_gpu_configs.py tests/test_gpu_edges.py -m gpu -q -x -k "not extreme_dims" > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
probe() { timeout 300 python - "$1" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
s = api.Session(x, g, D, J, K, N, 30, init_z=iz, rng_mode=1)
s.profile(True); s.run(); s.run()
st = s.stats(); n = st["n_iter"]
print(sys.argv[1], f"iters {n} loop {st['loop_ms']:.2f} ms: estep {st['family_ms']['estep']/n*1e3:.0f} us  partition {st['family_ms']['partition']/(n+1)*1e3:.0f} us  mcd {st['family_ms']['mcd']/(n+1)*1e3:.0f} us  rounds {st['mcd_passes']} incr {st['mcd_incr_rounds']} csteps {st['mcd_csteps']}")
PY
}
probe default
for v in variants/lib_*.so; do SQC_LIB=$PWD/$v probe $v; done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --rng counter > gpurun_out/bench_${TAG}_counter.json 2> gpurun_out/bench_${TAG}_counter.err
python - <<PY
import json
for m in ("", "_counter"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}%s.json" % m))
        r = d["roofline"]
        print(m, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % r["frac"], "mcd rounds", r["mcd_rounds_per_step"], "incr", r.get("mcd_incremental_rounds_per_step"), {k: round(v["avg_launch_ms"] * 1e3) for k, v in r["families"].items()})
    except Exception as e:
        print(m, "failed", e)
PY
timeout 300 python scripts/mcd_trace_all.py 1.0 30 1 > gpurun_out/${TAG}_trace_counter.log 2>&1; tail -3 gpurun_out/${TAG}_trace_counter.log
