#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/scorer_bench.py"
$CMD > gpurun_out/scorer_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 9 -c 1 -o gpurun_out/prof_scorer_single $CMD > gpurun_out/ncu_scorer.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_scorer.log
