#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/chain_timeline.py > gpurun_out/chain_timeline.log 2>&1; echo "rc=$?"; grep -A40 "iteration 3" gpurun_out/chain_timeline.log
