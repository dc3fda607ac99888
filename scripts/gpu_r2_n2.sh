#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== check_sharded"; timeout 600 $TR --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_n2.log 2>&1; echo "rc=$? t=$SECONDS"; grep -E "IDENTICAL|DIFFERENT|Error" gpurun_out/check_sharded_n2.log | head
for c in cfg3 cfg4 cfg5; do
  echo "== bench $c n2"; timeout 900 $TR --master-port 29541 bench.py --gpus 2 --config $c > gpurun_out/bench_${c}_n2.log 2> gpurun_out/bench_${c}_n2.err; echo "rc=$? t=$SECONDS"; grep -E "Error|error" gpurun_out/bench_${c}_n2.err | tail -3 | cut -c1-300; tail -1 gpurun_out/bench_${c}_n2.log | cut -c1-300
done
echo "== cfg5 n1 (parity band)"; timeout 600 python bench.py --config cfg5 > gpurun_out/bench_cfg5.log 2> gpurun_out/bench_cfg5.err; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/bench_cfg5.log | cut -c1-200
