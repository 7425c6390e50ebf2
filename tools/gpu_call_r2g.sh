#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 100 ./tools/microbench6 2>&1) > gpurun_out/r2g_microbench6.txt
(timeout 900 python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -60) > gpurun_out/r2g_gputests.txt
(timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3) > gpurun_out/r2g_bench.txt
tail -n 50 gpurun_out/r2g_*.txt
