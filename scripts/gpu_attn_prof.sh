#!/bin/bash
mkdir -p gpurun_out
export B=1024
CMD="python scripts/attn_bench.py"
$CMD > gpurun_out/attn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pim_attn_persistent -s 2 -c 1 -o gpurun_out/prof_attn_p1 $CMD > gpurun_out/attn_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/attn_plain.log; tail -2 gpurun_out/attn_ncu.log
