#!/usr/bin/env bash
# round 2, GPU call C: full parity suite with the split-K stepper as default + the K-grouped tensor-core weight gradients,
# gradient errors, bench line, ncu capture (full set + source) of fwd4s_kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -60) > gpurun_out/r2c_gputests.txt
(timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2c_graderr_tc.txt
(RNDE_BWD_FFMA=1 timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2c_graderr_ffma.txt
(timeout 300 python tools/bwd_ab.py 2>&1 | tail -4) > gpurun_out/r2c_bwd_ab_tc.txt
(timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3) > gpurun_out/r2c_bench.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd4s -s 2 -c 1 -f -o gpurun_out/prof_r2c_fwd4s python tools/fwd_time.py 2>&1 | tail -5) > gpurun_out/r2c_ncu.txt
tail -n 40 gpurun_out/r2c_*.txt
