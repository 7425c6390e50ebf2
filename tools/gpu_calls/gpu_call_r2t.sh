#!/usr/bin/env bash
# 2 GPUs: bench line with exact_check; 1 GPU part: the new A6 tests
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -3) > gpurun_out/r2t_bench_n2.txt
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "first_dt" 2>&1 | tail -5) > gpurun_out/r2t_a6_tests.txt
tail -c 900 gpurun_out/r2t_bench_n2.txt; cat gpurun_out/r2t_a6_tests.txt
