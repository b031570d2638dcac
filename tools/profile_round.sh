#!/bin/bash
# Round profile: (1) ncu launch list of exactly one CFG+DDIM step, (2) ncu --set full of the dominant kernel as bench.py times it
# (32x32-level epipolar attention with the tile map), (3) ncu --set full of representative GEMM / norm / dense-attention launches.
set -x
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
tail -1 gpurun_out/one_step.log
ncu --set full --clock-control none --import-source on -k regex:attn_fa_kernel -s 3 -c 1 -f -o gpurun_out/prof_epi0_map python tools/kernel_bench.py single epi0map > gpurun_out/prof_epi0_map.log 2>&1
tail -1 gpurun_out/prof_epi0_map.log
for spec in "lin0 gemm_tc_kernel" "conv0 gemm_tc_kernel" "geglu0 gemm_tc_kernel" "gn0 gn_apply_kernel" "attn0 attn_fa_kernel"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_$1 python tools/kernel_bench.py single $1 > gpurun_out/prof_$1.log 2>&1
  tail -1 gpurun_out/prof_$1.log
done
ls -la gpurun_out/*.ncu-rep
