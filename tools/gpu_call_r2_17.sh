#!/bin/bash
# Round 2, GPU call 17: mma.sync attention for 16 < n <= 64 tokens (the spatial transformer's frames).
set -u
O=gpurun_out/r2c17
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_models.py tests/test_gpu_decode.py tests/test_gpu_kernels.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; grep -i "attention" $O/op_bench.txt
for mid in 1 0; do
timeout 600 env PHK_MID_ATTN_MMA=$mid python bench.py --no-cpu --no-refgpu --no-train --no-maskgit > $O/bench_mid$mid.json 2> $O/bench_mid$mid.err; tail -c 200 $O/bench_mid$mid.err
python - "$O/bench_mid$mid.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "roofline", d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("step_frac_of_tensor_peak"))
    v = d.get("make_video")
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host enqueue", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
NCU="ncu --clock-control none --cache-control none"
PHK_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_encode_bf16.csv python tools/profile_step.py encode bf16 3 > $O/p_encode.log 2>&1
python tools/last_step.py $O/launches_encode_bf16.csv encode > $O/launches_encode_bf16.txt 2>&1; head -14 $O/launches_encode_bf16.txt
