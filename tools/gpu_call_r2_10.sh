#!/bin/bash
# Round 2, GPU call 10: the critic / primed demasking iteration as one replayed launch sequence (make_video path).
set -u
O=gpurun_out/r2c10
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_decode.py tests/test_gpu_masks_and_self_critic.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_fused_qkv.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -6 $O/tests.log
timeout 600 python bench.py --no-cpu --no-refgpu --no-train > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 600 env PHK_STEP_GRAPH=0 python bench.py --no-cpu --no-refgpu --no-train > $O/bench_loop.json 2> $O/bench_loop.err
for f in bench bench_loop; do
python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    v = d.get("make_video")
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host enqueue", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
