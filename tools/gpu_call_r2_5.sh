#!/bin/bash
# Round 2, GPU call 5: training step with batched attention-backward products + PEG dw reduction; LN fusion off (A/B on);
# primed fused sampling step; cosine-VQ at K = 65536; fresh launch lists.
set -u
O=gpurun_out/r2c5
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_masks_and_self_critic.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_parity_at_size.py tests/test_gpu_decode.py -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -4 $O/tests.log
timeout 600 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.err
timeout 600 env PHK_FUSE_LN=1 python bench.py --no-cpu --no-refgpu --no-train --no-makevideo > $O/bench_fuse_ln.json 2> $O/bench_fuse_ln.err
for f in bench bench_fuse_ln; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("  train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
    if v: print("  make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s", v.get("error", ""))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
NCU="ncu --clock-control none --cache-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train_bf16.csv python tools/train_bench.py 4 1 bf16 > $O/p_train.log 2>&1
python - $O/launches_train_bf16.csv <<'PY' > $O/launches_train_bf16.txt 2>&1
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
names = [re.sub(r'\(.*', '', r['Kernel Name']).replace('void phk::', '').replace('phk::', '') for r in rows]
idx = [i for i, nm in enumerate(names) if 'token_embed_kernel' in nm][-1]
agg, tot = {}, 0.0
for r, nm in list(zip(rows, names))[idx:]:
    v = float(r['Metric Value'].replace(',', '')) / 1000
    k = (nm[:60], r['Grid Size'])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print('last training step (bf16, b=4): sum of kernel durations', round(tot, 1), 'us')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {a[1]/tot*100:5.1f}% {k[0]:60s} grid={k[1]:18s} n={a[0]:3d} avg={a[1]/a[0]:9.1f} us")
PY
head -24 $O/launches_train_bf16.txt
for w in encode maskgit; do
  PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
  cat $O/launches_${w}_bf16.txt
done
