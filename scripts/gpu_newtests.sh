#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -k "skips_negative" > gpurun_out/new_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/new_pytest.log | cut -c1-250
