#!/bin/bash
# round-2 evidence pass on one B200: compute-sanitizer over the new kernels' tests, ncu captures of the scorer family's other
# modes (log-sum-exp, rank, top-k pass 1) and of the CE backward inside their bench configurations
mkdir -p gpurun_out
echo "== memcheck (new kernels)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q --no-header -x -p no:cacheprovider \
  -k "topk_tensor_core or exclusions_update or sharded_rank or d256 or few_rows or probability_dropout or returns_lse" > gpurun_out/memcheck_r2.log 2>&1
echo "memcheck rc=$? t=$SECONDS"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r2.log | tail -3
cap() {  # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  echo "ncu $name rc=$? t=$SECONDS"
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.csv 2>/dev/null
  python scripts/ncu_hot.py gpurun_out/$name.ncu-rep 30 > gpurun_out/${name}_hot.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
CMD5="python bench.py --config cfg5 --steps 1 --warmup 3 --no-parity"
# launch order of score_tc_kernel in one cfg5 step: <1> lse (get_pp), <2> rank x2 (rr increase), <3> top-k pass 1 (Caser)
cap r2_prof_lse "score_tc_kernel" 9 $CMD5
cap r2_prof_rank "score_tc_kernel" 10 $CMD5
cap r2_prof_topk_pass1 "score_tc_kernel" 11 $CMD5
CMD4="python bench.py --config cfg4 --steps 1 --warmup 3 --no-parity --no-cpu-baseline"
cap r2_prof_ce_bwd "ce_bwd_tc_kernel" 6 $CMD4
grep -h "Kernel Name" -A2 gpurun_out/r2_prof_{lse,rank,topk_pass1,ce_bwd}_summary.csv | cut -c1-120
