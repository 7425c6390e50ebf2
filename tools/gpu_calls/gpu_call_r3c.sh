#!/usr/bin/env bash
# 2 GPUs with the final library of round 2: exact-mode check, the exact-mode pytest, the N=2 bench line and its reference arm
mkdir -p gpurun_out
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_exact_check.py 512 2>&1 | grep -v Warning | tail -9) > gpurun_out/r3c_exact_check_n2.txt
(timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "exact_data_parallel" 2>&1 | tail -5) > gpurun_out/r3c_exact_pytest.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -2) > gpurun_out/r3c_bench_n2.txt
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300) > gpurun_out/r3c_ref_n2.txt
tail -n 12 gpurun_out/r3c_*.txt
