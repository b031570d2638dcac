#!/bin/bash
# End-of-round evidence in one GPU call: GPU test suite, smoke, default bench line, launch list of one step, --set full capture of the
# dominant kernel, warm in-graph durations by shape.  Outputs under gpurun_out/<tag>_*; summarise with tools/summarize_profiles.py.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/${tag}_gpu_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2 > gpurun_out/${tag}_smoke.txt
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/${tag}_bench_line.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_step.csv python tools/one_step.py > gpurun_out/${tag}_one_step.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fa_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_prof_epi0_map python tools/kernel_bench.py single epi0map > gpurun_out/${tag}_prof_epi0_map.log 2>&1
timeout 300 python tools/graph_trace.py --serial-passes --by-shape > gpurun_out/${tag}_graph_trace_by_shape.txt 2>&1
cat gpurun_out/${tag}_gpu_suite.txt gpurun_out/${tag}_smoke.txt gpurun_out/${tag}_bench_line.json
