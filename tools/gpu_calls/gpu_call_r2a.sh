#!/usr/bin/env bash
# round 2, GPU call A: side-branch verification, suite, microbench4, gradient errors of both sweeps, baseline bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
(RNDE_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_ffjord.py -x -q 2>&1 | tail -15) > gpurun_out/r2a_ffjord.txt
(timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -40) > gpurun_out/r2a_gputests.txt
(timeout 120 ./tools/microbench4 2>&1) > gpurun_out/r2a_microbench4.txt
(timeout 300 python tools/grad_err.py 512 2>&1 | tail -4) > gpurun_out/r2a_graderr_tc.txt
(RNDE_BWD_FFMA=1 timeout 300 python tools/grad_err.py 512 2>&1 | tail -4) > gpurun_out/r2a_graderr_ffma.txt
(timeout 300 python tools/grad_err.py 32 2>&1 | tail -4) >> gpurun_out/r2a_graderr_tc.txt
(RNDE_BWD_FFMA=1 timeout 300 python tools/grad_err.py 32 2>&1 | tail -4) >> gpurun_out/r2a_graderr_ffma.txt
(timeout 300 python tools/bwd_ab.py 2>&1 | tail -4) > gpurun_out/r2a_bwd_ab_tc.txt
(RNDE_BWD_FFMA=1 timeout 300 python tools/bwd_ab.py 2>&1 | tail -4) > gpurun_out/r2a_bwd_ab_ffma.txt
(timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2a_bench.txt
tail -n 30 gpurun_out/r2a_*.txt
