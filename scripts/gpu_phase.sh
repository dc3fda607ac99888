#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -k "two_phase or argmax" > gpurun_out/phase_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/phase_pytest.log | cut -c1-300
