#!/bin/bash
# Round 2, GPU call 28: the whole GPU suite + smoke() + bench line on the head of the branch.
set -u
O=gpurun_out/r2c28
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_gpu.log 2>&1; echo "gpu suite exit=$?"; tail -3 $O/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 $O/smoke.log
timeout 600 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | e2e", d.get("e2e", {}).get("value"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step")
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
    if v: print("make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print("unreadable:", ex)
PY
