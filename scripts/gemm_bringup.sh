#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -k "linear_tensor_core or argmax_tensor_core" > gpurun_out/gemm_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/gemm_pytest.log
timeout 200 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "bench rc=$?"; tail -12 gpurun_out/gemm_bench.log
