#!/bin/bash
# Round 2, GPU call 4: temporal attention on mma.sync, bench with the configs[3] / configs[4] blocks, launch list of
# the training step.
set -u
O=gpurun_out/r2c4
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_qkv.py tests/test_gpu_x3_mode.py tests/test_gpu_bf16_mode.py tests/test_gpu_fullsize.py tests/test_gpu_parity_at_size.py tests/test_gpu_gemm_bf16.py tests/test_gpu_models.py tests/test_gpu_decode.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
timeout 600 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    print("train_step", json.dumps(d.get("train_step"))[:700])
    print("make_video", json.dumps(d.get("make_video"))[:900])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train_bf16.csv python tools/train_bench.py 4 1 bf16 > $O/p_train.log 2>&1
python - $O/launches_train_bf16.csv <<'PY' > $O/launches_train_bf16.txt 2>&1
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void phk::', '').replace('phk::', '') for r in rows]
# the last training step = launches after the last token_embed_kernel
idx = [i for i, nm in enumerate(names) if 'token_embed_kernel' in nm][-1]
agg, tot = {}, 0.0
for r, nm in list(zip(rows, names))[idx:]:
    v = float(r['Metric Value'].replace(',', '')) / 1000
    k = (nm[:60], r['Grid Size'])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print('last training step (bf16, b=4): sum of kernel durations', round(tot, 1), 'us')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {a[1]/tot*100:5.1f}% {k[0]:60s} grid={k[1]:18s} n={a[0]:3d} avg={a[1]/a[0]:9.1f} us")
PY
cat $O/launches_train_bf16.txt | head -45
PHK_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_encode_bf16.csv python tools/profile_step.py encode bf16 3 > $O/p_encode.log 2>&1
python tools/last_step.py $O/launches_encode_bf16.csv encode | head -16
