#!/bin/bash
# End-of-round evidence in one GPU call: GPU test suite, smoke, default bench line + reference arm, launch list of one step,
# --set full captures of the dominant kernels, warm in-graph durations by shape (with and without programmatic dependent launch:
# PDL overlaps a kernel's launch with its predecessor, which inflates the per-kernel durations CUPTI reports), launch lists of the
# once-per-sample stages.  Outputs under gpurun_out/<tag>_*; summarise with tools/summarize_profiles.py and
# tools/ncu_top_stalls.py.   REF_STEPS / REF_WARMUP: size of the reference-arm run (default 5 / 2).
tag=${1:-final}
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -q -m gpu -s 2>&1 | grep -E "passed|failed|FAILED|rel-L2|vs pure|loop vs|tile map" | tail -40 > gpurun_out/${tag}_gpu_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2 > gpurun_out/${tag}_smoke.txt
timeout 400 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_line.json
timeout 600 python bench.py --impl reference --steps ${REF_STEPS:-5} --warmup ${REF_WARMUP:-2} 2>/dev/null | tail -1 > gpurun_out/${tag}_reference_line.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_step.csv python tools/one_step.py > gpurun_out/${tag}_one_step.log 2>&1
for spec in "epi0map attn_fa_kernel" "attn0 attn_fa_kernel" "lin0 gemm_tc_kernel" "lincat0 gemm_tc_kernel" "conv0 gemm_tc_kernel" "geglu0 gemm_ps_kernel" "qkv0 gemm_ps_kernel" "gn0 gn_apply_kernel"; do
  set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/${tag}_prof_$1 python tools/kernel_bench.py single $1 > gpurun_out/${tag}_prof_$1.log 2>&1
done
timeout 300 python tools/graph_trace.py --serial-passes --by-shape > gpurun_out/${tag}_graph_trace_by_shape.txt 2>&1
C2V_PDL=0 timeout 300 python tools/graph_trace.py --serial-passes --by-shape > gpurun_out/${tag}_graph_trace_by_shape_pdl0.txt 2>&1
timeout 300 python tools/kernel_bench.py > gpurun_out/${tag}_kernel_bench.txt 2>&1
for f in "adaptor " "resampler --resampler" "vae --vae" "pose --pose-encoder"; do
  set -- $f
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$1.csv python tools/adaptor_bench.py $2 --no-cpu > gpurun_out/${tag}_frow_$1.log 2>&1
done
cat gpurun_out/${tag}_gpu_suite.txt gpurun_out/${tag}_smoke.txt gpurun_out/${tag}_bench_line.json gpurun_out/${tag}_reference_line.json
