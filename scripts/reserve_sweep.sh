for r in 20 22 24 26 28; do
SQC_MCD_RESERVE=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/bench_sw_res$r.json 2> /dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/bench_sw_res$r.json")); rf = d["roofline"]
print("reserve", $r, "value %.4g" % d["value"], "ms/step %.2f" % d["ms_per_step"], {k: round(v["avg_launch_ms"] * 1e3) for k, v in rf["families"].items()})
PY
done
