#!/bin/bash
# Round 2, GPU call 22: CTA-scope "accumulator drained" arrive in the CTA-pair GEMM (no ERRBAR in front of it).
set -u
O=gpurun_out/r2c22
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm_bf16.py tests/test_gpu_fused_qkv.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_train.py tests/test_gpu_decode.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; grep -i "gemm\|head\|tail" $O/op_bench.txt | head -24
timeout 600 python bench.py --no-cpu --no-refgpu --no-makevideo > $O/bench.json 2> $O/bench.err; tail -c 200 $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t = d.get("train_step")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
