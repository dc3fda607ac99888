#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_train_dp.py > gpurun_out/train_dp2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/train_dp2.log | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/bench_train_dp.py --per-gpu-batch 4096 > gpurun_out/train_dp2w.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/train_dp2w.log | cut -c1-500
