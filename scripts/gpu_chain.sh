#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -x -k "decoder_chain" > gpurun_out/chain_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/chain_pytest.log | cut -c1-250
timeout 200 python scripts/chain_bench.py > gpurun_out/chain_bench.log 2>&1; echo "bench rc=$?"; tail -6 gpurun_out/chain_bench.log
timeout 200 python scripts/chain_timeline.py > gpurun_out/chain_timeline.log 2>&1; echo "rc=$?"; grep -A45 "iteration 3" gpurun_out/chain_timeline.log
