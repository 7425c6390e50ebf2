#!/usr/bin/env bash
# 2 GPUs: reference-exact mode with the first-dt term (sums over all ranks' columns), the exact-mode pytest, the N=2 bench line
mkdir -p gpurun_out
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_exact_check.py 512 2>&1 | grep -v Warning | tail -9) > gpurun_out/r2z_exact_check_n2.txt
(timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "exact_data_parallel" 2>&1 | tail -5) > gpurun_out/r2z_exact_pytest.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -2) > gpurun_out/r2z_bench_n2.txt
tail -n 12 gpurun_out/r2z_*.txt
