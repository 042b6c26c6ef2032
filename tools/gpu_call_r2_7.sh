#!/bin/bash
# Round 2, GPU call 7: cross-attention on mma.sync; LayerNorm in the residual GEMM epilogue with the statistics exchanged
# through global memory (PHK_FUSE_LN=2); the whole GPU suite with split-bf16 as the default precision.
set -u
O=gpurun_out/r2c7
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_masks_and_self_critic.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -4 $O/tests.log
timeout 900 env PHK_PREC=bf16x3 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_default_x3.log 2>&1; echo "x3-default suite exit=$?"; tail -12 $O/tests_default_x3.log
timeout 600 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s", v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
timeout 600 env PHK_FUSE_LN=2 python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_fuse_ln2.json 2> $O/bench_fuse_ln2.err; tail -c 300 $O/bench_fuse_ln2.err
python - "$O/bench_fuse_ln2.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
NCU="ncu --clock-control none --cache-control none"
PHK_FUSE_LN=2 PHK_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_encode_fuse_ln2.csv python tools/profile_step.py encode bf16 3 > $O/p_encode_ln2.log 2>&1
python tools/last_step.py $O/launches_encode_fuse_ln2.csv encode > $O/launches_encode_fuse_ln2.txt 2>&1; head -12 $O/launches_encode_fuse_ln2.txt
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_maskgit_bf16.csv python tools/profile_step.py maskgit bf16 3 > $O/p_maskgit.log 2>&1
python tools/last_step.py $O/launches_maskgit_bf16.csv maskgit > $O/launches_maskgit_bf16.txt 2>&1; head -9 $O/launches_maskgit_bf16.txt
NCUF="ncu --clock-control none --set full --import-source on"
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 300 $NCUF -k regex:"attention_tc_kernel|head_sample_kernel|attention_cross_mma|gemm_bf16_kernel|gemm_bf16_pair" -s 30 -c 16 -o $O/maskgit_full -f python tools/profile_step.py maskgit bf16 2 > $O/p_maskgit_full.log 2>&1
python tools/ncu_summary.py $O/maskgit_full.ncu-rep $O/maskgit_full_summary.csv; ls -la $O/maskgit_full.ncu-rep
cut -c1-420 $O/maskgit_full_summary.csv
