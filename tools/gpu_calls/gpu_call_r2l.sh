#!/bin/bash
# GPU call L: first-dt gradient term (row A6) -- new test, then the gradient suites, then the bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "first_dt" 2>&1 | tail -30 > gpurun_out/r2l_a6.txt
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2l_suite.txt
python __graft_entry__.py smoke > gpurun_out/r2l_smoke.txt 2>&1
python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -5 gpurun_out/r2l_a6.txt; tail -8 gpurun_out/r2l_suite.txt; tail -3 gpurun_out/r2l_smoke.txt; cat gpurun_out/r2l_bench.json
