#!/bin/bash
# dense_wgrad_kernel with doubles staged in shared memory: gradient tests of every caller (FFJORD, Latent ODE chain + GRU, toy chain), then the timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ffjord.py tests/test_gpu_latent.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2zy_tests.txt
cat gpurun_out/r2zy_tests.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2zy_times.txt
import sys, json, torch
sys.path.insert(0, ".")
import bench
import numpy as np
import regneuralde.jl_b200 as R
out = bench.secondary_workloads(torch, R, 71.98)
print(json.dumps(out, indent=1))
PY
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_wgrad --csv python tools/ffjord_step.py 2>&1 | tail -4) | tee -a gpurun_out/r2zy_times.txt
