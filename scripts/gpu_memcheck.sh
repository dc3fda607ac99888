#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_kernels.py -q --no-header -p no:cacheprovider -x -k "decoder_chain_tensor_core or writes_operand_images or (pim_attention_forward and 201) or (softmax_ce and 260) or (lse and 130) or (argmax_tensor_core_equals and 130)" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/memcheck.log | cut -c1-200
