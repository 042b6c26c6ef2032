#!/bin/bash
# Round 2, GPU call 1: parity of the benchmarked modes vs the reference's own CUDA paths, the same-box GPU bar,
# the one-graph-launch-per-iteration A/B, first GPU run of the training step, fresh launch lists.
set -u
O=gpurun_out/r2c1
mkdir -p $O
timeout 420 python tools/parity_report.py $O/parity.json > $O/parity.log 2>&1; echo "parity exit=$?"
tail -5 $O/parity.log
timeout 240 python tools/ref_gpu_bench.py > $O/ref_gpu.json 2> $O/ref_gpu.err; echo "ref_gpu exit=$?"
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 env PHK_STEP_GRAPH=1 python bench.py --no-cpu > $O/bench_step_graph.json 2> $O/bench_step_graph.err
for f in default step_graph; do python - "$O/bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("extra", {}).get("maskgit_sample", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
timeout 200 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; tail -8 $O/op_bench.txt
export PHK_EXPERIMENTAL=1
for c in with_critic self_critic generator; do
  timeout 200 python tests/gpu_train_check.py $c > $O/train_check_$c.log 2>&1; echo "exit=$?" >> $O/train_check_$c.log
  tail -2 $O/train_check_$c.log
done
for c in generator with_critic; do
  timeout 200 python tests/gpu_train_check.py --bf16 $c > $O/train_check_bf16_$c.log 2>&1; echo "exit=$?" >> $O/train_check_bf16_$c.log
  tail -2 $O/train_check_bf16_$c.log
done
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/gpu_train_check.py with_critic > $O/train_memcheck.log 2>&1; echo "exit=$?" >> $O/train_memcheck.log
tail -3 $O/train_memcheck.log
timeout 300 python tools/train_bench.py 4 3 f32 > $O/train_bench_f32.json 2> $O/train_bench_f32.err; cat $O/train_bench_f32.json
timeout 300 python tools/train_bench.py 4 5 bf16 > $O/train_bench_bf16.json 2> $O/train_bench_bf16.err; cat $O/train_bench_bf16.json
NCU="ncu --clock-control none --cache-control none"
for w in encode maskgit; do
  PHK_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
done
ls -la $O
