#!/bin/bash
# DESIGN section 9 in ONE gpurun call (each call costs ~2.5 GPU-minutes of overhead):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Output: gpurun_out/first/ -- the GPU suite, the bench line three ways (default / all-rows head / one graph launch per
# demasking iteration), the per-op times incl. the masked-rows tail, the training-step validation, fresh launch lists.
set -u
O=gpurun_out/first
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "exit=$?" >> $O/gpu_tests.log
tail -4 $O/gpu_tests.log
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 env PHK_HEAD_COMPACT=0 python bench.py --no-cpu > $O/bench_allrows_head.json 2> $O/bench_allrows_head.err
timeout 300 env PHK_STEP_GRAPH=1 python bench.py --no-cpu > $O/bench_step_graph.json 2> $O/bench_step_graph.err
for f in default allrows_head step_graph; do python - "$O/bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("extra", {}).get("maskgit_sample", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; tail -8 $O/op_bench.txt
bash tools/validate_train_gpu.sh > $O/train.log 2>&1; tail -12 $O/train.log
NCU="ncu --clock-control none --cache-control none"
for w in encode maskgit; do
  PHK_GRAPH=0 timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
done
ls -la $O
