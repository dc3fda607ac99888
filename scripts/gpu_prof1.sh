#!/bin/bash
# Launch list (per-kernel device time) of one short cfg3 bench, then an ncu --set full capture of the scorer.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-path-len 2"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/prof_plain.log
$CMD > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_simt -s 3 -c 2 -o gpurun_out/prof_score_simt_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
