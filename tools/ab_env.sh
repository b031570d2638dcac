#!/bin/bash
# Interleaved A/B of a host-side switch (an environment variable read by camc2v_b200) on ONE box:
#     gpurun --timeout 600 -- 'bash tools/ab_env.sh C2V_TT_FUSE 0 1'
# $ROUNDS rounds (default 2) of the default bench (no CPU baseline) per value, interleaved.  Output: gpurun_out/ab_env.txt
cd "$(dirname "$0")/.."
var=$1; shift
out=gpurun_out/ab_env.txt; mkdir -p gpurun_out; : > $out
for r in $(seq ${ROUNDS:-2}); do for v in "$@"; do
  env $var=$v timeout 100 python bench.py --no-cpu-baseline 2>&1 | tail -1 | grep -o "\"ms_per_step\": [0-9.]*\|\"sm_mhz\": [0-9.]*" | grep -v ': $' | tr '\n' ' ' | sed "s/^/$var=$v /" >> $out
  echo >> $out
done; done
cat $out
