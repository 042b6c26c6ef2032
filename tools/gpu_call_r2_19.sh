#!/bin/bash
# Round 2, GPU call 19: LayerNorm fold, stage 1 (feed-forward and cross-attention LayerNorms inside the products).
# Kept for provenance: it ran at commit a205ec7; the code it exercises (PHK_LN_FOLD, phk_gemm_bf16_res_stats, ..._fold) was
# reverted by fa35a21 after these measurements (profiles/r02/*_c19_*).
set -u
O=gpurun_out/r2c19
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_gemm_bf16.py tests/test_gpu_bf16_mode.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -8 $O/tests.log
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; grep -i "gemm" $O/op_bench.txt
for f in 0 1; do
timeout 600 env PHK_LN_FOLD=$f python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_fold$f.json 2> $O/bench_fold$f.err; tail -c 200 $O/bench_fold$f.err
python - "$O/bench_fold$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
timeout 600 env PHK_LN_FOLD=1 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_models.py tests/test_gpu_decode.py -q -p no:cacheprovider > $O/tests_fold_on.log 2>&1; echo "suite with fold on exit=$?"; tail -6 $O/tests_fold_on.log
cp gpurun_out/parity_at_size.json $O/parity_at_size_fold_on.json 2>/dev/null
