#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/bench_secondary.py cfg2 cfg5 cfg4 > gpurun_out/bench_secondary.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_secondary.log | cut -c1-900
