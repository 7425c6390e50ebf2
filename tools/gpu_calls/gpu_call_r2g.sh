#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 100 ./tools/microbench6 2>&1) > gpurun_out/r2g_microbench6.txt
(timeout 300 python -m pytest tests/test_gpu_nsde.py -q -x 2>&1 | tail -40) > gpurun_out/r2g_nsde.txt
(timeout 900 python -m pytest tests -q -m gpu -s --deselect tests/test_gpu_nsde.py 2>&1 | grep -v "^$" | tail -60) > gpurun_out/r2g_gputests.txt
(timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3) > gpurun_out/r2g_bench.txt
tail -n 50 gpurun_out/r2g_*.txt
