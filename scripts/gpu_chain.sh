#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -x -k "decoder_chain or chain" > gpurun_out/chain_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/chain_pytest.log | cut -c1-250
timeout 200 python scripts/chain_bench.py 2>&1 | tee gpurun_out/chain_bench.log
timeout 200 python scripts/chain_timeline.py > gpurun_out/chain_timeline.log 2>&1; echo "rc=$?"; grep -A60 "iteration 3" gpurun_out/chain_timeline.log | grep -B100 "iteration 4" | grep -E "EPI  (wait_d1|d|e1|e3|e4)|MMA  (a0_ready|g1_issued|a1_ready|ffn_issued|a3_ready|g4)"
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropins.py -q --no-header -rf -p no:cacheprovider -x > gpurun_out/model_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/model_pytest.log | cut -c1-250
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [ (k[:12], round(v['ms_per_launch'],4)) for k,v in d['roofline_kernels'].items()])"
