#!/bin/bash
# Round 2, GPU call 14: cross-attention on packed operands; ncu --set full of the fused logits head.
set -u
O=gpurun_out/r2c14
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_models.py tests/test_gpu_decode.py tests/test_gpu_masks_and_self_critic.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
for pk in 1 0; do
timeout 600 env PHK_CROSS_PACK=$pk python bench.py --no-cpu --no-refgpu --no-train > $O/bench_pack$pk.json 2> $O/bench_pack$pk.err; tail -c 200 $O/bench_pack$pk.err
python - "$O/bench_pack$pk.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step e2e", (m.get("e2e") or {}).get("value"), m.get("error", ""))
    v = d.get("make_video")
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host enqueue", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
NCU="ncu --clock-control none --cache-control none"
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_maskgit_bf16.csv python tools/profile_step.py maskgit bf16 3 > $O/p_maskgit.log 2>&1
python tools/last_step.py $O/launches_maskgit_bf16.csv maskgit > $O/launches_maskgit_bf16.txt 2>&1; head -14 $O/launches_maskgit_bf16.txt
NCUF="ncu --clock-control none --set full --import-source on"
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 400 $NCUF -k regex:"head_sample_kernel|attention_cross_packed|cross_kv_pack" -s 3 -c 6 -o $O/head_full -f python tools/profile_step.py maskgit bf16 2 > $O/p_head_full.log 2>&1
python tools/ncu_summary.py $O/head_full.ncu-rep $O/head_full_summary.csv && rm -f $O/head_full.ncu-rep
cut -c1-330 $O/head_full_summary.csv
