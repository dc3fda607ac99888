#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -x -p no:cacheprovider -k "argmax or topk or rank or lse" > gpurun_out/pytest_scorer.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_scorer.log | cut -c1-300
timeout 300 python scripts/scorer_bench2.py > gpurun_out/scorer_bench2.log 2>&1; echo "rc=$?"; cat gpurun_out/scorer_bench2.log
timeout 300 python scripts/scorer_timeline.py > gpurun_out/scorer_timeline.log 2>&1; tail -12 gpurun_out/scorer_timeline.log
