#!/bin/bash
# r2 profiling pass: phase timelines, ncu launch list, one `ncu --set full` capture per hot kernel; summaries (raw metrics +
# hottest SASS lines) are produced on the box, the big .ncu-rep files are dropped (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
echo "== timelines"
timeout 300 python scripts/chain_timeline.py > gpurun_out/chain_timeline.log 2>&1; echo "chain rc=$?"
timeout 300 python scripts/attn_timeline.py > gpurun_out/attn_timeline.log 2>&1; echo "attn rc=$?"
timeout 300 python scripts/scorer_timeline.py > gpurun_out/scorer_timeline.log 2>&1; echo "scorer rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-parity --e2e-path-len 2"
echo "== plain"; $CMD > gpurun_out/prof_plain.log 2>&1; echo "rc=$? t=$SECONDS"
echo "== launch list"; ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "rc=$? t=$SECONDS"
cap() {  # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  echo "ncu $name rc=$? t=$SECONDS"
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.csv 2>/dev/null
  python scripts/ncu_hot.py gpurun_out/$name.ncu-rep 40 > gpurun_out/${name}_hot.txt 2>/dev/null
}
cap r2_prof_chain "decoder_chain_kernel" 20 $CMD
cap r2_prof_attn "pim_attn_persistent" 20 $CMD
cap r2_prof_scorer "score_tc_kernel" 3 $CMD
cap r2_prof_rescore "rescore_finalize" 3 $CMD
cap r2_prof_gather "embed_gather" 3 $CMD
CMD5="python bench.py --config cfg5 --steps 1 --warmup 3 --no-parity"
cap r2_prof_topk_pass1 "score_tc_kernel<3>|score_tc_kernelILi3" 1 $CMD5
cap r2_prof_topk_select "topk_select" 1 $CMD5
cap r2_prof_rank "score_tc_kernel<2>|score_tc_kernelILi2" 2 $CMD5
rm -f gpurun_out/r2_prof_rescore.ncu-rep gpurun_out/r2_prof_gather.ncu-rep gpurun_out/r2_prof_topk_pass1.ncu-rep gpurun_out/r2_prof_topk_select.ncu-rep gpurun_out/r2_prof_rank.ncu-rep gpurun_out/r2_prof_scorer.ncu-rep
du -sh gpurun_out
