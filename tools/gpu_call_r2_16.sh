#!/bin/bash
# Round 2, GPU call 16: A/B data -- GEMM tile variants on the current code, programmatic dependent launch on / off,
# CUDA-graph replay on / off.
set -u
O=gpurun_out/r2c16
mkdir -p $O
for mode in 1 2 3; do
  echo "PHK_GEMM_MODE=$mode"; PHK_GEMM_MODE=$mode timeout 200 python tools/op_bench.py 50 2>&1 | head -7 | tee $O/op_bench_mode$mode.txt
done
for cfg in "default" "PHK_PDL=0" "PHK_GRAPH=0 PHK_STEP_GRAPH=0" "PHK_PDL=0 PHK_GRAPH=0 PHK_STEP_GRAPH=0"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  if [ "$cfg" = "default" ]; then envs=""; else envs="$cfg"; fi
  timeout 400 env $envs python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_$tag.json 2> $O/bench_$tag.err
  python - "$O/bench_$tag.json" "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(f"{sys.argv[2]:45s} encode {d['ms_per_step']:.4f} ms | sustained {d.get('sustained', {}).get('ms_per_step')} | maskgit {m.get('ms_per_decode_step')} ms/step e2e {(m.get('e2e') or {}).get('value')}")
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
