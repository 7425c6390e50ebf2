#!/bin/bash
# dense_wgrad_kernel build variants (staging batch, CTAs per SM, threads per CTA): duration on the FFJORD training step
# (record of what ran: build_variants/lib_*.so were whole-library builds of intermediate revisions of chain.cuh with -DRNDE_CW_BATCH / _MINB / _NT / _TPT;
#  the directory is not kept -- results in profiles/r2zz_next_rows_ncu.txt)
mkdir -p gpurun_out; : > gpurun_out/r2zx_variants.txt
for v in b1 b4 b8m2 b4m2 b4t512 b2t512m2; do
  cp build_variants/lib_$v.so regneuralde/jl_b200/libregnde.so
  echo "== $v" >> gpurun_out/r2zx_variants.txt
  (timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_wgrad --csv python tools/ffjord_step.py 2>&1 | tail -2 | awk -F, '{print $NF}') >> gpurun_out/r2zx_variants.txt
done
cat gpurun_out/r2zx_variants.txt
