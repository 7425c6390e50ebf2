#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" | tail -30) > gpurun_out/r2k_gputests.txt
(timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1) > gpurun_out/r2k_bench.txt
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4) > gpurun_out/r2k_smoke.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd4s -s 12 -c 1 -f -o gpurun_out/prof_r2k_fwd4s_taped python tools/fwd_time.py 2>&1 | tail -3) > gpurun_out/r2k_ncu.txt
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1)
tail -n 30 gpurun_out/r2k_gputests.txt gpurun_out/r2k_smoke.txt gpurun_out/r2k_ncu.txt; cut -c1-300 gpurun_out/r2k_bench.txt
