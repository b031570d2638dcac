#!/bin/bash
# GPU side of tools/ab_build.sh: interleaved A/B of the variant libraries on ONE box (box-to-box variance is +-0.3 ms per step,
# so variants must be compared inside one gpurun call).
#     gpurun --timeout 900 -- 'bash tools/ab_run.sh base maskpf biaspf'
# Per variant: the kernel micro-benchmarks named in $SINGLES, then $ROUNDS rounds of the default bench (no CPU baseline), then the
# kernel tests once per non-base variant.  Output: gpurun_out/ab.txt
cd "$(dirname "$0")/.."
out=gpurun_out/ab.txt; mkdir -p gpurun_out; : > $out
SINGLES=${SINGLES:-"epi0map attn0 lin0 geglu0"}
ROUNDS=${ROUNDS:-2}
for k in "$@"; do
  export CAMC2V_B200_LIB=$PWD/variants/$k.so
  echo "== $k" >> $out
  for m in $SINGLES; do timeout 60 python tools/kernel_bench.py single $m 2>&1 | tail -1 >> $out; done
done
for r in $(seq $ROUNDS); do for k in "$@"; do
  export CAMC2V_B200_LIB=$PWD/variants/$k.so
  timeout 100 python bench.py --no-cpu-baseline 2>&1 | tail -1 | grep -o "\"ms_per_step\": [0-9.]*" | grep -v ': $' | sed "s/^/$k /" >> $out
done; done
for k in "$@"; do
  [ "$k" = base ] && continue
  export CAMC2V_B200_LIB=$PWD/variants/$k.so
  echo "== tests $k" >> $out
  timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -2 >> $out
done
cat $out
