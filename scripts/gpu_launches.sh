#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-path-len 2"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
grep -E "score_tc_kernel|rescore_finalize" gpurun_out/launches_final.csv | awk -F'","' '{print $5, $NF}' | cut -c1-50,190- | tail -8
