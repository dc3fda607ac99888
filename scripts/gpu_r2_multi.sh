#!/bin/bash
# N=<2|4|8> bash scripts/gpu_r2_multi.sh : sharded-vs-single identity + cfg3 / cfg4 / cfg5 bench lines at N GPUs
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== check_sharded n$N"; timeout 240 $TR --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_n$N.log 2>&1; echo "rc=$? t=$SECONDS"; grep -E "IDENTICAL|DIFFERENT|Error" gpurun_out/check_sharded_n$N.log | head -4
for c in ${CONFIGS:-cfg3 cfg4 cfg5}; do
  echo "== bench $c n$N"; timeout 300 $TR --master-port 29541 bench.py --gpus $N --config $c > gpurun_out/bench_${c}_n$N.log 2> gpurun_out/bench_${c}_n$N.err; echo "rc=$? t=$SECONDS"; grep -E "Error|error" gpurun_out/bench_${c}_n$N.err | tail -3 | cut -c1-300; tail -1 gpurun_out/bench_${c}_n$N.log | cut -c1-250
done
