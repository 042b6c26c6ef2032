#!/bin/bash
# Round 2, GPU call 27: PEG stencil on packed FFMA2.
set -u
O=gpurun_out/r2c27
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -4 $O/tests.log
timeout 300 python tools/op_bench.py 50 2>&1 | grep -i "peg" 
timeout 400 python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench.json 2> $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step")
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
