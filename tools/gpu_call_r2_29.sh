#!/bin/bash
# Round 2, GPU call 29: the in-situ test of the mid-size attention path.
set -u
O=gpurun_out/r2c29
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bf16_mode.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -6 $O/tests.log
