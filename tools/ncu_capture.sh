#!/bin/bash
# ncu captures of the top kernels (run under gpurun; one GPU).  Usage: tools/ncu_capture.sh <tag> [family ...]
# Writes gpurun_out/<tag>_<family>.ncu-rep (+ raw CSV page) and keeps gpurun_out/ under the 64 MiB merge limit.
set -u
tag=$1; shift
fams=${@:-"ew reduce permute topk gemm attn"}
mkdir -p gpurun_out
declare -A KRE=( [ew]="ew_" [reduce]="reduce_" [permute]="transpose_" [topk]="topk_" [gemm]="gemm_tc2" [attn]="attn_fwd_tc" [attn_bwd]="attn_bwd" [gemm_f32]="gemm_f32x|split_f32" [norm]="moments|layer_norm" [attn_f32]="attn_f32_tc|split_planes" [attn_pers]="attn_fwd_pers" )
declare -A CNT=( [ew]=2 [reduce]=5 [permute]=1 [topk]=2 [gemm]=1 [attn]=1 [attn_bwd]=3 [gemm_f32]=3 [norm]=4 [attn_f32]=4 [attn_pers]=1 )
for f in $fams; do
  out=gpurun_out/${tag}_${f}
  KF_PROF_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KRE[$f]} -c ${CNT[$f]} -f -o $out \
      python tools/prof_ops.py $f > $out.log 2>&1
  if [ -f $out.ncu-rep ]; then
    ncu -i $out.ncu-rep --page raw --csv > $out.raw.csv 2>/dev/null
    sz=$(stat -c %s $out.ncu-rep)
    if [ $sz -gt 12000000 ]; then ncu -i $out.ncu-rep --page source --csv > $out.source.csv 2>/dev/null; rm -f $out.ncu-rep; fi
  fi
done
du -sh gpurun_out; ls -la gpurun_out
