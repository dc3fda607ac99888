#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q --no-header -rf -p no:cacheprovider -k "cfg3_shape or full_size" > gpurun_out/new_pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/new_pytest.log | cut -c1-250
