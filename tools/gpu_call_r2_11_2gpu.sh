#!/bin/bash
# Round 2, 2-GPU call: data-parallel training step (overlapped sliced all-reduce vs whole-bucket), bench at N = 2.
set -u
O=gpurun_out/r2c11
mkdir -p $O
timeout 300 python bench.py --no-cpu --no-refgpu --no-train --no-maskgit --steps 5 --warmup 3 > $O/bench_mv.json 2> $O/bench_mv.err
python - "$O/bench_mv.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    v = d.get("make_video")
    if v: print("N=1 make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host enqueue", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py > $O/ddp_check.log 2>&1; echo "ddp_check exit=$?"; grep DDP_CHECK $O/ddp_check.log || tail -20 $O/ddp_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench exit=$?"; tail -c 400 $O/bench_2gpu.err
python - "$O/bench_2gpu.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("N=2 encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", json.dumps(t.get("all_reduce")), t.get("error", ""))
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s", v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
