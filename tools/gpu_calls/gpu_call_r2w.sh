#!/bin/bash
# fresh ncu --set full captures of the final build: the tensor-core sweep with the first-dt additions, and one weight-gradient GEMM
mkdir -p gpurun_out
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwd4tc -s 2 -c 1 -f -o gpurun_out/prof_r2w_bwd4tc python tools/bwd_ab.py 2>&1 | tail -3) > gpurun_out/r2w_ncu.txt
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2w_wgrad_tc python tools/bwd_ab.py 2>&1 | tail -3) >> gpurun_out/r2w_ncu.txt
ls -la gpurun_out/*.ncu-rep; tail -8 gpurun_out/r2w_ncu.txt
