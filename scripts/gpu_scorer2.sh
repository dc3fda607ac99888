#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/scorer_bench.py"
$CMD > gpurun_out/scorer_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_scorer.csv $CMD > gpurun_out/ncu_scorer.log 2>&1
echo "rc=$?"; cat gpurun_out/scorer_plain.log; grep -E "score_tc_kernel|rescore" gpurun_out/launches_scorer.csv | awk -F'","' '{print $5, $NF}' | tail -12
