#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log | cut -c1-300
echo "== smoke" ; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for c in cfg3 cfg1 cfg2 cfg4 cfg5 k2; do
  echo "== bench $c"; timeout 900 python bench.py --config $c > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "rc=$? t=$SECONDS"; tail -2 gpurun_out/bench_$c.err | cut -c1-300; tail -1 gpurun_out/bench_$c.log | cut -c1-400
done
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$? t=$SECONDS"; tail -2 gpurun_out/bench_ref.err | cut -c1-300; tail -1 gpurun_out/bench_ref.log | cut -c1-400
