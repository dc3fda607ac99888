#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-300
for c in ${CONFIGS:-cfg3}; do
  echo "== bench $c"; timeout 900 python bench.py --config $c ${BENCH_ARGS:---no-cpu-baseline --no-gpu-baseline} > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "rc=$? t=$SECONDS"; tail -2 gpurun_out/bench_$c.err | cut -c1-300; tail -1 gpurun_out/bench_$c.log | cut -c1-200
done
