#!/bin/bash
# Round 2, 2-GPU call: overlapped sliced gradient all-reduce (zero-element parameters no longer void the plan).
set -u
O=gpurun_out/r2c12
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py > $O/ddp_check.log 2>&1; echo "ddp_check exit=$?"; grep DDP_CHECK $O/ddp_check.log || tail -20 $O/ddp_check.log
for ov in 1 0; do
PHK_OVERLAP_ALL_REDUCE=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-refgpu --no-maskgit --no-makevideo > $O/bench_2gpu_ov$ov.json 2> $O/bench_2gpu_ov$ov.err; echo "bench exit=$?"; tail -c 300 $O/bench_2gpu_ov$ov.err
python - "$O/bench_2gpu_ov$ov.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t = d.get("train_step")
    if t: print(sys.argv[1], "train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", json.dumps(t.get("all_reduce")), t.get("overlap"), t.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
