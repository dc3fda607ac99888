#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; echo "rc=$?"; head -120 gpurun_out/attn_timeline.log
