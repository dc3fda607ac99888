#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropins.py tests/test_gpu_model.py -q --no-header -rf -p no:cacheprovider -x -k "lse or softmax_ce or evaluator or train or loss" > gpurun_out/lse_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/lse_pytest.log | cut -c1-250
timeout 1200 python scripts/bench_secondary.py cfg4 > gpurun_out/bench_secondary4.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bench_secondary4.log | cut -c1-900
