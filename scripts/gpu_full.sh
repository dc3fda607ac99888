#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log | cut -c1-250
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== chain bench"; timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/chain_bench.log
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_cfg3.log | cut -c1-900
