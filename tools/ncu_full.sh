#!/bin/bash
# One `ncu --set full` capture per representative kernel (skip the 3 warm-up launches), reports under gpurun_out/.
mkdir -p gpurun_out
for spec in "lin0 gemm_tc_kernel" "conv0 gemm_tc_kernel" "gn0 gn_apply_kernel" "epi0 attn_fa_kernel" "attn0 attn_fa_kernel"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_$1 python tools/kernel_bench.py single $1 > gpurun_out/prof_$1.log 2>&1
  tail -1 gpurun_out/prof_$1.log
done
ls -la gpurun_out/*.ncu-rep
