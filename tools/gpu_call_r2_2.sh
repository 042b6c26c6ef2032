#!/bin/bash
# Round 2, GPU call 2: fused QKV operand path (GEMM epilogue normalisation, strided attention_tc, MN-major V, temporal bf16)
set -u
O=gpurun_out/r2c2
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_gemm_bf16.py -x -q -p no:cacheprovider > $O/tests_kernels.log 2>&1; echo "kernel tests exit=$?"; tail -5 $O/tests_kernels.log
timeout 600 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_decode.py tests/test_gpu_parity_at_size.py tests/test_gpu_masks_and_self_critic.py tests/test_gpu_train.py -x -q -p no:cacheprovider > $O/tests_models.log 2>&1; echo "model tests exit=$?"; tail -5 $O/tests_models.log
timeout 300 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err
timeout 300 env PHK_FUSE_QKV=0 python bench.py --no-cpu --no-refgpu > $O/bench_nofuse.json 2> $O/bench_nofuse.err
for f in bench bench_nofuse; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""), "| e2e", m.get("e2e", {}).get("value"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
NCU="ncu --clock-control none --cache-control none"
for w in encode maskgit; do
  PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
  cat $O/launches_${w}_bf16.txt
done
