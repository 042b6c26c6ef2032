#!/bin/bash
# Round 2, GPU call 23: source-level ncu captures (one launch each) of the demasking iteration's kernels.
set -u
O=gpurun_out/r2c23
mkdir -p $O
NCUF="ncu --clock-control none --set full --import-source on"
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 500 $NCUF -k regex:"gemm_bf16_kernel|attention_tc_kernel|head_sample_kernel|peg_tiled|attention_cross_packed|ln_warp_kernel|gemm_bf16_pair" -s 40 -c 24 -o $O/maskgit_src -f python tools/profile_step.py maskgit bf16 2 > $O/p.log 2>&1
ls -la $O
