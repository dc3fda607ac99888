#!/bin/bash
mkdir -p gpurun_out
for s in 0 1; do
timeout 200 python scripts/chain_timeline.py $s > gpurun_out/chain_timeline_$s.log 2>&1; echo "stagger/skip=$s rc=$?"; grep -A40 "iteration 3" gpurun_out/chain_timeline_$s.log | grep -E "EPI  (wait_d1|d1_ready|e1_done|d3_ready|e3_done|d4ab_ready|e4ab_done|d4c_ready|e4c_done)"
done
