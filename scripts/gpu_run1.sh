#!/bin/bash
# First GPU contact: smoke, parity tests, a short bench.  Everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
echo "== bench small"; timeout 300 python bench.py --small --steps 5 --users 512 > gpurun_out/bench_small.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_small.log
echo "== bench cfg3"; timeout 900 python bench.py --steps 5 > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_cfg3.log
