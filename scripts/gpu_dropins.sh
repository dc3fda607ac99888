#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dropins.py -q --no-header -rf -p no:cacheprovider > gpurun_out/dropins_pytest.log 2>&1
echo "pytest rc=$?"; tail -60 gpurun_out/dropins_pytest.log | cut -c1-250
