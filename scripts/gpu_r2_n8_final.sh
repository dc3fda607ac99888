#!/bin/bash
# 8 GPUs, short timeouts: the whole cfg3 job (100k users x 20 steps) sharded over the box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== soak n8"; timeout 150 $TR --master-port 29543 scripts/soak_cfg3.py > gpurun_out/soak_n8.log 2>&1; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/soak_n8.log | cut -c1-300
