#!/bin/bash
# ncu --set full of the two small per-iteration kernels of the MLE variant (per-sample reduction + alpha solve; moment reduction + update)
mkdir -p gpurun_out
for k in k_mle_sample k_mle_reduce; do
  SCALE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o /tmp/ncu_$k -f python scripts/mle_bench.py C3 6 > gpurun_out/ncu_$k.log 2>&1
  ncu -i /tmp/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$k.ncu-rep --page source --csv > gpurun_out/ncu_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out/ncu_k_mle_* | cut -c1-200
