#!/bin/bash
# Round 2, GPU call 21: source-level ncu capture of the FF1 + GEGLU CTA-pair kernel (where do the membar / barrier stalls sit?)
set -u
O=gpurun_out/r2c21
mkdir -p $O
NCUF="ncu --clock-control none --set full --import-source on"
PHK_GRAPH=0 timeout 400 $NCUF -k regex:"gemm_bf16_pair_kernel" -s 8 -c 1 -o $O/pair_full -f python tools/profile_step.py encode bf16 2 > $O/p_pair_full.log 2>&1
ls -la $O
