#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -x -k "pim_attention_forward" > gpurun_out/attn_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/attn_pytest.log | cut -c1-220
timeout 200 python scripts/attn_bench.py > gpurun_out/attn_bench.log 2>&1; echo "bench rc=$?"; tail -6 gpurun_out/attn_bench.log
timeout 200 python scripts/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; echo "rc=$?"; sed -n 30,75p gpurun_out/attn_timeline.log
