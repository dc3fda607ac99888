#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "rank" -q --no-header -rf -p no:cacheprovider > gpurun_out/rank_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/rank_pytest.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_dropins.py tests/test_gpu_model.py -q --no-header -rf -p no:cacheprovider > gpurun_out/dropins_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/dropins_pytest.log | cut -c1-300
timeout 600 python scripts/bench_secondary.py cfg5 > gpurun_out/bench_secondary_rank.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_secondary_rank.log | cut -c1-700
