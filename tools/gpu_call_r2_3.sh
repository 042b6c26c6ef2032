#!/bin/bash
# Round 2, GPU call 3: reworked in-kernel noise (Philox-7, log2-domain gumbel) in the fused head; ncu --set full of the
# kernels that dominate the two steps on the current code.
set -u
O=gpurun_out/r2c3
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_bf16_mode.py tests/test_gpu_kernels.py tests/test_gpu_zz_after_last_gpu_call.py tests/test_gpu_models.py tests/test_gpu_x3_mode.py tests/test_gpu_parity_at_size.py tests/test_gpu_fused_qkv.py -x -q -p no:cacheprovider > $O/tests.log 2>&1; echo "tests exit=$?"; tail -5 $O/tests.log
timeout 300 python bench.py --no-cpu --no-refgpu > $O/bench.json 2> $O/bench.err; tail -c 400 $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d.get("maskgit", {})
    print(sys.argv[1], "encode", round(d["value"]), "frames/s", d["ms_per_step"], "ms | sustained", d.get("sustained", {}).get("ms_per_step"), "| maskgit", m.get("value"), "tokens/s", m.get("ms_per_decode_step"), "ms/step", m.get("error", ""), "| e2e", m.get("e2e", {}).get("value"), "| launches/iter", m.get("kernels_per_iteration"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
timeout 300 python bench.py --prec bf16x3 --no-cpu --no-refgpu --no-maskgit > $O/bench_x3.json 2> $O/bench_x3.err; python -c "
import json; d=json.loads(open('$O/bench_x3.json').read().strip().splitlines()[-1]); print('bf16x3 encode', round(d['value']), 'frames/s', d['ms_per_step'], 'ms')"
timeout 300 python bench.py --prec f32 --no-cpu --no-refgpu --no-maskgit > $O/bench_f32.json 2> $O/bench_f32.err; python -c "
import json; d=json.loads(open('$O/bench_f32.json').read().strip().splitlines()[-1]); print('f32 encode', round(d['value']), 'frames/s', d['ms_per_step'], 'ms')"
timeout 200 python tools/op_bench.py 50 > $O/op_bench.txt 2>&1; tail -9 $O/op_bench.txt
NCU="ncu --clock-control none --set full --import-source on"
PHK_GRAPH=0 PHK_STEP_GRAPH=0 timeout 300 $NCU -k regex:"attention_tc_kernel|head_sample_kernel|attention_fewkeys|gemm_bf16_kernel|peg_tiled" -s 20 -c 14 -o $O/maskgit_full -f python tools/profile_step.py maskgit bf16 2 > $O/p_maskgit_full.log 2>&1
PHK_GRAPH=0 timeout 300 $NCU -k regex:"attention_warp64|attention_tc_kernel|gemm_bf16_kernel|patchify|ln_warp|ln_lfq|peg_tiled|gemm_bf16_pair" -s 4 -c 24 -o $O/encode_full -f python tools/profile_step.py encode bf16 1 > $O/p_encode_full.log 2>&1
for r in maskgit_full encode_full; do
  python tools/ncu_summary.py $O/$r.ncu-rep $O/${r}_summary.csv; ls -la $O/$r.ncu-rep
done
cat $O/maskgit_full_summary.csv | cut -c1-400
