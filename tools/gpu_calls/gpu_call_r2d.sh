#!/usr/bin/env bash
# round 2, GPU call D: software-pipelined layer loops in fwd4s, prefetching wgrad loader; the 1.5x gradient bars
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -70) > gpurun_out/r2d_gputests.txt
(timeout 200 python tools/fwd_time.py 2>&1 | tail -3) > gpurun_out/r2d_fwd_time.txt
(timeout 300 python tools/fwd4_timeline.py tape 2>&1 | tail -25) > gpurun_out/r2d_timeline.txt
(timeout 300 python tools/bwd_ab.py 2>&1 | tail -4) > gpurun_out/r2d_bwd_ab_tc.txt
(timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2d_bench.txt
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2d_wgrad python tools/bwd_ab.py 2>&1 | tail -3) > gpurun_out/r2d_ncu.txt
tail -n 45 gpurun_out/r2d_*.txt
