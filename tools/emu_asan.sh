#!/bin/bash
# "compute-sanitizer memcheck" without a GPU: the plain-CUDA sources compiled by g++ for the CPU executor (tests/cuda_emu)
# with -fsanitize=address, libasan preloaded into python, and the emulated suites run on top -- every out-of-bounds
# access of a kernel to a torch buffer or to the workspace allocation is reported with the kernel's source line.
# ~25 min on 8 cores.  usage: bash tools/emu_asan.sh [pytest args]
set -u
export PHK_EMU_ASAN=1
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0
exec python -m pytest tests/test_kernels_emulated_cpu.py tests/test_models_emulated_cpu.py tests/test_train_emulated_cpu.py \
  tests/test_sample_tail_emulated_cpu.py tests/test_bf16_drivers_emulated_cpu.py -x -q -p no:cacheprovider "$@"
