#!/bin/bash
# round-2 verification on one B200: GPU tests, smoke, every bench configuration, both reference arms, the whole-job soak
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke" ; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in cfg3 cfg1 cfg2 cfg4 cfg5 k2; do
  echo "== bench $c"; timeout 600 python bench.py --config $c > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/bench_$c.log | cut -c1-160
done
for v in zipf ragged; do
  echo "== bench cfg3 $v"; timeout 600 python bench.py --variant $v --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_cfg3_$v.log 2> gpurun_out/bench_cfg3_$v.err; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/bench_cfg3_$v.log | cut -c1-160
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/bench_ref.log | cut -c1-160
echo "== torch_gpu arm"; timeout 600 python bench.py --impl torch_gpu > gpurun_out/bench_torch_gpu.log 2> gpurun_out/bench_torch_gpu.err; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/bench_torch_gpu.log | cut -c1-300
echo "== soak"; timeout 300 python scripts/soak_cfg3.py > gpurun_out/soak_n1.log 2>&1; echo "rc=$? t=$SECONDS"; tail -1 gpurun_out/soak_n1.log | cut -c1-300
