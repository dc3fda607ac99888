#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-path-len 2"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -1 gpurun_out/prof_plain.log | cut -c1-200
