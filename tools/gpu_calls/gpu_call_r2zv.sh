#!/bin/bash
# dense_wgrad_kernel: CTAs per SM and layer (RNDE_CW_OVERSUB) on the FFJORD training step
mkdir -p gpurun_out; : > gpurun_out/r2zv_oversub.txt
for o in 2 3 4 6 8; do
  echo "== oversub $o" >> gpurun_out/r2zv_oversub.txt
  (RNDE_CW_OVERSUB=$o timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_wgrad --csv python tools/ffjord_step.py 2>&1 | tail -2 | awk -F, '{print $(NF-6), $NF}') >> gpurun_out/r2zv_oversub.txt
done
cat gpurun_out/r2zv_oversub.txt
