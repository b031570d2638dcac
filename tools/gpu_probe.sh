#!/bin/bash
# Run every kernel-level GPU test function in its own process (a trapped kernel poisons the CUDA context),
# with a hard timeout each; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
python -c "import oracle; oracle.build()" 
TESTS=${@:-test_linear test_linear_strided test_geglu test_conv3x3 test_conv_t3 test_skinny test_groupnorm test_layernorm test_attention_dense test_attention_epipolar test_attention_temporal test_epipolar_mask_bit_exact test_plucker test_layout_and_glue test_downsample test_cfg_ddim_update}
for t in $TESTS; do
  echo "=== $t" | tee -a gpurun_out/probe.log
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$t" --timeout 500 -x 2>&1 | tail -25 | tee -a gpurun_out/probe.log
done
grep -E "^===|passed|failed|error" gpurun_out/probe.log | tail -60
