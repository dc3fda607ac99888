#!/bin/bash
mkdir -p gpurun_out
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 > gpurun_out/bench_cfg3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_cfg3.log | cut -c1-3000
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-path-len 2"
$CMD > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"pim_attn_persistent|decoder_chain_kernel" -s 14 -c 3 -o gpurun_out/prof_r1_attn_chain $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
