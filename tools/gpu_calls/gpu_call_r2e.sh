#!/usr/bin/env bash
# round 2, GPU call E: loop-bottleneck microbenchmarks, backward-sweep timeline, fwd timing of the reverted (unrolled) loop
mkdir -p gpurun_out
(timeout 120 ./tools/microbench5 2>&1) > gpurun_out/r2e_microbench5.txt
(timeout 200 python tools/fwd_time.py 2>&1 | tail -3) > gpurun_out/r2e_fwd_time.txt
(timeout 300 python tools/bwd4tc_timeline.py 2>&1 | tail -20) > gpurun_out/r2e_bwd_timeline.txt
tail -n 45 gpurun_out/r2e_*.txt
