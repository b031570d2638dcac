#!/bin/bash
# Build A/B variants of the default (fp16-operand) library with extra nvcc defines, here on the CPU box:
#     tools/ab_build.sh base: maskpf:-DC2V_AT_MASK_PREFETCH=1 biaspf:-DC2V_GEMM_BIAS_PREFETCH=1
# writes variants/<name>.so (git-ignored, but it travels to the GPU box with gpurun); restores the default build at the end.
# Run them with tools/ab_run.sh on the GPU box.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  echo "== $name ($defs)"
  C2V_NVCC_EXTRA="$defs" python -m camc2v_b200.build --fp16 --force 2>&1 | grep -i "error" || true
  cp camc2v_b200/libcamc2v_b200_fp16.so variants/$name.so
done
python -m camc2v_b200.build --fp16 --force 2>&1 | grep -i "error" || true
ls -la variants/
