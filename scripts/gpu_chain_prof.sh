#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/chain_bench.py 205824"
$CMD > gpurun_out/chain_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decoder_chain -s 3 -c 1 -o gpurun_out/prof_chain_v3 $CMD > gpurun_out/chain_ncu.log 2>&1
echo "ncu rc=$?"; tail -4 gpurun_out/chain_plain.log; tail -2 gpurun_out/chain_ncu.log
