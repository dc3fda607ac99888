#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q --no-header -rf -p no:cacheprovider -x -k "argmax or generat or full_size or paths or lse or softmax_ce" > gpurun_out/scorer_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/scorer_pytest.log | cut -c1-250
python scripts/scorer_bench.py
