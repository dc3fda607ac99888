#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log | cut -c1-300
echo "== smoke" ; timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench cfg3"; timeout 600 python bench.py > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3.log | cut -c1-300
nproc; free -g | head -2
