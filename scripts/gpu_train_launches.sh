#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/bench_secondary.py cfg4"
$CMD > gpurun_out/train_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_train.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/train_plain.log | cut -c1-200
