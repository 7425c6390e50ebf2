#!/usr/bin/env bash
# 4 GPUs (the one world size not yet run in round 2): exact-mode check and the bench line
mkdir -p gpurun_out
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/dist_exact_check.py 512 2>&1 | grep -v Warning | tail -6) > gpurun_out/r3f_exact_check_n4.txt
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 2>&1 | grep '^{') > gpurun_out/r3f_bench_n4.json
cat gpurun_out/r3f_exact_check_n4.txt; cut -c1-400 gpurun_out/r3f_bench_n4.json
