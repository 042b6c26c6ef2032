#!/bin/bash
# Round 2, 8-GPU call: the whole bench line at N = 8 (encode, MaskGIT, configs[3] training step with the overlapped
# all-reduce, configs[4] make_video chains batch-sharded over the GPUs).
set -u
O=gpurun_out/r2c15
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-refgpu > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench exit=$?"; tail -c 400 $O/bench_8gpu.err
python - "$O/bench_8gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("N=8 encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | e2e", d.get("e2e", {}).get("value"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", json.dumps(t.get("all_reduce")), t.get("overlap"), t.get("error", ""))
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
nproc; free -g | head -2
