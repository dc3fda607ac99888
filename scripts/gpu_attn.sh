#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --no-header -rf -p no:cacheprovider -k "attention or operand_images" > gpurun_out/attn_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/attn_pytest.log | cut -c1-250
timeout 200 python scripts/attn_bench.py 2>&1 | head -1
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [ (k[:12], round(v['ms_per_launch'],4)) for k,v in d['roofline_kernels'].items()])"
