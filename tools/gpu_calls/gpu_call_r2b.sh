#!/usr/bin/env bash
# round 2, GPU call B: first run of the split-K stepper (fwd4s_kernel): parity suite, A/B timing against fwd4_kernel,
# microbench4, gradient errors with the round-to-nearest wgrad split
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -40) > gpurun_out/r2b_gputests.txt
(timeout 120 ./tools/microbench4 2>&1) > gpurun_out/r2b_microbench4.txt
(RNDE_ARITH=0 timeout 200 python tools/fwd_time.py 2>&1 | tail -3) > gpurun_out/r2b_fwd_time_arith0.txt
(RNDE_ARITH=2 timeout 200 python tools/fwd_time.py 2>&1 | tail -3) > gpurun_out/r2b_fwd_time_arith2.txt
(timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2b_graderr_tc.txt
(RNDE_BWD_FFMA=1 timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2b_graderr_ffma.txt
(RNDE_BWD_FFMA=1 RNDE_WGRAD_FFMA=1 timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2b_graderr_ffma_wgradffma.txt
(RNDE_WGRAD_FFMA=1 timeout 300 python tools/grad_err.py 512 2>&1 | tail -2) > gpurun_out/r2b_graderr_tc_wgradffma.txt
(timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2b_bench.txt
tail -n 30 gpurun_out/r2b_*.txt
