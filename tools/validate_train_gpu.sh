#!/bin/bash
# First GPU run of the training step.  Under gpurun:  bash tools/validate_train_gpu.sh
# Output: gpurun_out/train/ -- parity of loss + every gradient against the reference's autograd for the three golden
# cases, a compute-sanitizer memcheck pass over the smallest case, and the step time at BASELINE configs[2] sizes.
set -u
O=gpurun_out/train
mkdir -p $O
export PHK_EXPERIMENTAL=1
for c in with_critic self_critic generator; do
  timeout 300 python tests/gpu_train_check.py $c > $O/check_$c.log 2>&1; echo "exit=$?" >> $O/check_$c.log
  tail -3 $O/check_$c.log
done
for c in generator with_critic; do
  timeout 300 python tests/gpu_train_check.py --bf16 $c > $O/check_bf16_$c.log 2>&1; echo "exit=$?" >> $O/check_bf16_$c.log
  tail -3 $O/check_bf16_$c.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/gpu_train_check.py with_critic > $O/memcheck.log 2>&1; echo "exit=$?" >> $O/memcheck.log
tail -5 $O/memcheck.log
timeout 600 python tools/train_bench.py 4 3 f32 > $O/train_bench_f32.json 2> $O/train_bench_f32.err; cat $O/train_bench_f32.json
timeout 600 python tools/train_bench.py 4 5 bf16 > $O/train_bench_bf16.json 2> $O/train_bench_bf16.err; cat $O/train_bench_bf16.json
