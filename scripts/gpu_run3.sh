#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
echo "== bench cfg3"; timeout 900 python bench.py --steps 20 > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-path-len 2"
$CMD > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_tc_max -s 3 -c 1 -o gpurun_out/prof_score_tc_v2 $CMD > gpurun_out/ncu_full3.log 2>&1
echo "ncu rc=$?"
