#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -x -k "decoder_chain" > gpurun_out/chain_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/chain_pytest.log | cut -c1-250
timeout 200 python scripts/chain_timeline.py > gpurun_out/chain_timeline.log 2>&1; echo "rc=$?"; grep -A40 "iteration 3" gpurun_out/chain_timeline.log | grep -E "EPI|MMA  (a0_ready|g1_issued|a1_ready|ffn_issued|a3_ready|g4)"
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3.log | cut -c1-400
