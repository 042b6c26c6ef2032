#!/bin/bash
# Round 2, final capture on one B200: the whole GPU suite, smoke(), the full bench line (CPU baseline + same-box GPU bar),
# the reference arm, launch lists of every workload, ncu --set full of the GEMM family / attention / head, op bench.
# usage (under gpurun): bash tools/gpu_call_r2_final.sh [tag]   -> gpurun_out/r2final[tag]/
set -u
O=gpurun_out/r2final${1:-}
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/tests_gpu.log 2>&1; echo "gpu suite exit=$?"; tail -4 $O/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit=$?"; tail -c 300 $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm exit=$?"
python - "$O/bench.json" "$O/bench_reference.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print("encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| e2e", d.get("e2e", {}).get("value"), "| roofline", d.get("roofline", {}).get("frac"), "traffic", d.get("roofline", {}).get("traffic"))
    print("maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step e2e", (m.get("e2e") or {}).get("value"), "roofline", (m.get("roofline") or {}).get("frac"), m.get("error", ""))
    t, v = d.get("train_step"), d.get("make_video")
    if t: print("train_step", t.get("ms_per_step"), "ms", t.get("value"), "tokens/s", t.get("error", ""))
    if v: print("make_video", v.get("ms_per_chain"), "ms/chain", v.get("value"), "tokens/s host", v.get("host_enqueue_ms_per_chain"), v.get("error", ""))
    print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:300])
    print("reference_gpu", json.dumps(d.get("reference_gpu"))[:600])
    r = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("reference arm", r.get("value"), r.get("unit"), r.get("ms_per_step"), "ms/step", json.dumps(r.get("maskgit"))[:200])
except Exception as ex:
    print("unreadable:", ex)
PY
timeout 300 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1
NCU="ncu --clock-control none --cache-control none"
for w in encode decode maskgit; do
  PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
  python tools/last_step.py $O/launches_${w}_bf16.csv $w > $O/launches_${w}_bf16.txt 2>&1 || true
  head -5 $O/launches_${w}_bf16.txt
done
PHK_GRAPH=0 timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_train_bf16.csv python tools/train_bench.py 4 2 bf16 > $O/p_train.log 2>&1
python tools/last_step.py $O/launches_train_bf16.csv train > $O/launches_train_bf16.txt 2>&1 || true
NCUF="ncu --clock-control none --set full --import-source on"
# GEMM family of one encode step (second step: launches 61..): DRAM traffic per launch -> profiles/r02_traffic.json
PHK_GRAPH=0 timeout 400 $NCUF -k regex:"gemm_bf16" -s 33 -c 33 -o $O/gemm_full -f python tools/profile_step.py encode bf16 2 > $O/p_gemm_full.log 2>&1
PHK_GRAPH=0 timeout 400 $NCUF -k regex:"attention_tc_kernel|attention_small_mma|patchify_ln|peg_tiled|ln_lfq|ln_warp" -s 2 -c 20 -o $O/rest_full -f python tools/profile_step.py encode bf16 1 > $O/p_rest_full.log 2>&1
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 400 $NCUF -k regex:"head_sample_kernel|attention_tc_kernel|attention_cross_mma" -s 6 -c 4 -o $O/maskgit_full -f python tools/profile_step.py maskgit bf16 2 > $O/p_maskgit_full.log 2>&1
for r in gemm_full rest_full maskgit_full; do
  python tools/ncu_summary.py $O/$r.ncu-rep $O/${r}_summary.csv && rm -f $O/$r.ncu-rep
done
python - "$O/gemm_full_summary.csv" "$O/traffic.json" <<'PY'
import csv, json, sys
rows = list(csv.DictReader(open(sys.argv[1])))
mb = lambda s: float(s.split()[0]) if s else 0.0
n = sum(int(r["launches_captured"]) for r in rows)
tot = sum(mb(r["dram_rd_wr_mean"]) * int(r["launches_captured"]) for r in rows)
json.dump({"gemm_bf16": {"dram_bytes_per_launch": tot / max(n, 1) * 1e6, "source": f"profiles/r02/gemm_ncu_full_summary_final.csv: dram__bytes_read.sum + dram__bytes_write.sum averaged over the {n} GEMM launches of one encode step, ncu --set full --cache-control none"}}, open(sys.argv[2], "w"), indent=1)
print(open(sys.argv[2]).read())
PY
cut -c1-260 $O/gemm_full_summary.csv; cut -c1-260 $O/maskgit_full_summary.csv
ls -la $O | head -40
