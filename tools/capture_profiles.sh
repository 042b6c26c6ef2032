#!/bin/bash
# Runs on the GPU box (under gpurun): everything profiles/ is built from.  Output: gpurun_out/r01/
# usage: bash tools/capture_profiles.sh
set -u
O=gpurun_out/r01
mkdir -p $O
NCU="ncu --clock-control none --cache-control none"
python bench.py > $O/bench_bf16.json 2> $O/bench_bf16.err
python bench.py --prec f32 --no-cpu > $O/bench_f32.json 2> $O/bench_f32.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python tools/op_bench.py 50 > $O/op_bench.txt 2>&1
for w in encode decode maskgit; do
  PHK_GRAPH=0 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_${w}_bf16.csv python tools/profile_step.py $w bf16 3 > $O/p_$w.log 2>&1
done
# one --set full capture of the dominant kernel family (GEMM): 16 launches from the middle of the second encode step
PHK_GRAPH=0 $NCU --set full --import-source on -k regex:gemm_bf16 -s 36 -c 16 -o $O/gemm_full -f python tools/profile_step.py encode bf16 2 > $O/p_gemm_full.log 2>&1
# the other kernels of one encode step (first step: position-bias kernels included, so skip 2 launches)
PHK_GRAPH=0 $NCU --set full --import-source on -k regex:"attention_tc_kernel|attention_prep|attention_warp64|patchify_ln|peg_tiled|ln_lfq|ln_warp" -s 2 -c 24 -o $O/rest_full -f python tools/profile_step.py encode bf16 1 > $O/p_rest_full.log 2>&1
$NCU --set full --import-source on -k regex:"head_sample_kernel|attention_tc_kernel|attention_fewkeys" -s 6 -c 3 -o $O/maskgit_full -f python tools/profile_step.py maskgit bf16 2 > $O/p_maskgit_full.log 2>&1
# summaries only: the .ncu-rep files together exceed what gpurun copies back (64 MiB)
for r in gemm_full rest_full maskgit_full; do
  python tools/ncu_summary.py $O/$r.ncu-rep $O/${r}_summary.csv && rm -f $O/$r.ncu-rep
done
python tools/gemm_trace.py 4608 2816 512 2 > $O/gemm_trace_ff1.txt 2>&1
python tools/gemm_trace.py 4608 512 512 0 > $O/gemm_trace_qproj.txt 2>&1
ls -la $O
