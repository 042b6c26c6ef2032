#!/bin/bash
# Round 2, GPU call 20: GEGLU epilogue with one MUFU op per output (tanh.approx) against ex2 + rcp -- A/B of two builds.
set -u
O=gpurun_out/r2c20
mkdir -p $O
V=$PWD/phenaki_pytorch_b200/libphk_geglu_tanh.so
for lib in default tanh; do
  if [ $lib = tanh ]; then export PHK_LIB=$V; else unset PHK_LIB; fi
  timeout 200 python tools/op_bench.py 50 2>&1 | grep -i "FF1\|FF2" | sed "s/^/$lib: /"
  timeout 400 python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_$lib.json 2> $O/bench_$lib.err
  python - "$O/bench_$lib.json" $lib <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(f"{sys.argv[2]:8s} encode {d['ms_per_step']:.4f} ms | sustained {d.get('sustained', {}).get('ms_per_step')} | maskgit {m.get('ms_per_decode_step')} ms/step")
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
export PHK_LIB=$V
timeout 600 python -m pytest tests/test_gpu_gemm_bf16.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py -q -p no:cacheprovider > $O/tests_tanh.log 2>&1; echo "tests with the tanh build exit=$?"; tail -6 $O/tests_tanh.log
cp gpurun_out/parity_at_size.json $O/parity_at_size_tanh.json 2>/dev/null
