#!/usr/bin/env bash
mkdir -p gpurun_out
(timeout 300 python tools/sde_debug.py 2>&1 | tail -40) > gpurun_out/r2i_sde_debug.txt
(timeout 300 python tools/grad_blocks.py 2>&1 | tail -20) > gpurun_out/r2i_grad_blocks.txt
(timeout 300 python -m pytest tests/test_gpu_nsde.py -q 2>&1 | tail -15) > gpurun_out/r2i_nsde.txt
tail -n 45 gpurun_out/r2i_*.txt
