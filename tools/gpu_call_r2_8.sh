#!/bin/bash
# Round 2, GPU call 8: attention core with the probabilities in tensor memory (3 CTAs/SM); W-tile prefetch ahead of the
# dependency wait in the one-CTA GEMM; cluster LayerNorm epilogue fix; cross-attention mma test; make_video host share.
set -u
O=gpurun_out/r2c8
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_gemm_bf16.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_models.py tests/test_gpu_kernels.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -6 $O/tests.log
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; cat $O/op_bench.txt
timeout 600 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host enqueue", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
timeout 300 env PHK_ATTN_P_TMEM=0 python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_p_smem.json 2> $O/bench_p_smem.err
python - "$O/bench_p_smem.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
NCU="ncu --clock-control none --cache-control none"
for w in encode maskgit; do
  PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
  head -14 $O/launches_${w}_bf16.txt
done
