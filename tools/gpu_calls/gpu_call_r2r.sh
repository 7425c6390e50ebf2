#!/usr/bin/env bash
# 8 GPUs: reference-exact mode with the first-dt term -- bit-identity / gradient check, then the N=8 bench line
mkdir -p gpurun_out
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dist_exact_check.py 512 2>&1 | grep -v Warning | tail -9) > gpurun_out/r2r_exact_check_n8.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | tail -2) > gpurun_out/r2r_bench_n8.txt
tail -n 12 gpurun_out/r2r_exact_check_n8.txt; tail -c 600 gpurun_out/r2r_bench_n8.txt
