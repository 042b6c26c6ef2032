#!/bin/bash
# Round 2, GPU call 25: compute-sanitizer memcheck over the kernels added in the second half of the round
# (TMEM-probabilities attention, packed cross-attention, mid-size MMA attention, critic / primed iteration).
set -u
O=gpurun_out/r2c25
mkdir -p $O
CS="compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20"
timeout 1200 $CS python -m pytest tests/test_gpu_fused_qkv.py -q -p no:cacheprovider -x -k "packed or mid_mma or (attention_tc and (3-130 or 1-200 or 5-199)) or small_attention or cross_attention_bf16" > $O/memcheck_kernels.log 2>&1; echo "memcheck kernels exit=$?"; tail -6 $O/memcheck_kernels.log
timeout 1200 $CS python -m pytest tests/test_gpu_bf16_mode.py -q -p no:cacheprovider -x -k "critic_and_primed or one_graph_launch" > $O/memcheck_iterations.log 2>&1; echo "memcheck iterations exit=$?"; tail -6 $O/memcheck_iterations.log
grep -c "ERROR SUMMARY: 0 errors" $O/memcheck_kernels.log $O/memcheck_iterations.log
