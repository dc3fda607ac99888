#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.log 2>&1
echo "N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-330
done
