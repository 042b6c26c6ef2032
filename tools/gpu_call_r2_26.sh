#!/bin/bash
# Round 2, GPU call 26: attention-backward contractions (dP, dq, dk, dv) on warp-level tensor-core MMAs in bf16 training mode.
set -u
O=gpurun_out/r2c26
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_train.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
for mma in 1 0; do
  PHK_ATTN_BWD_MMA=$mma timeout 300 python tools/train_bench.py 4 8 bf16 > $O/train_bench_mma$mma.json 2> $O/train_bench_mma$mma.err; echo "PHK_ATTN_BWD_MMA=$mma"; cat $O/train_bench_mma$mma.json | cut -c1-400
done
timeout 300 python bench.py --no-cpu --no-refgpu --no-maskgit --no-makevideo > $O/bench.json 2> $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    t = d.get("train_step")
    if t: print("bench train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s loss", t.get("loss"), t.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
NCU="ncu --clock-control none --cache-control none"
PHK_GRAPH=0 timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train_bf16.csv python tools/train_bench.py 4 2 bf16 > $O/p_train.log 2>&1
python tools/last_step.py $O/launches_train_bf16.csv train > $O/launches_train_bf16.txt 2>&1; head -14 $O/launches_train_bf16.txt
