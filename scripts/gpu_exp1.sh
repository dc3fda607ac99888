#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -k "rank or argmax or score or lse or softmax" > gpurun_out/scorer_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/scorer_pytest.log | cut -c1-250
timeout 200 python scripts/scorer_bench.py 2>&1 | tee gpurun_out/scorer_plain.log
