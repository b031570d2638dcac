#!/bin/bash
# BASELINE.json configs[3] / configs[4] evidence on N GPUs of one box (run under `gpurun --gpus N`):
#     bash tools/multi_gpu_round.sh N tag [variants...]
#   * global batch 32 sharded by video (32 / N videos per GPU, no step-time collective, final NCCL all_gather of the latents)
#   * the same with the CFG halves split over rank pairs (one 2-rank all_gather of the noise predictions per step)
#   * optional: the baseline conditioning variants (cami2v / cameractrl / motionctrl) at global batch 32
# One JSON line per run under gpurun_out/<tag>_n<N>_*.json (value = video-steps/s of the whole job, videos_per_s = value / 25).
N=${1:-2}; tag=${2:-r02}; shift 2
mkdir -p gpurun_out
run() {   # name, extra flags
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
      --steps 5 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/${tag}_n${N}_${name}.err | tail -1 > gpurun_out/${tag}_n${N}_${name}.json
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${N}_${name}.json").read())
    print("${name}", "N=${N}", "steps/s", round(d["value"], 2), "videos/s", round(d["videos_per_s"], 3), "ms/step", round(d["ms_per_step"], 2),
          "videos/GPU", d["config"]["videos_per_gpu"], "per-rank ms", [round(x, 1) for x in d["per_rank"]["ms_per_step"]])
except Exception as e:
    print("${name}", "FAILED", e)
PY
}
run gb32 --global-batch 32
run gb32_cfgsplit --global-batch 32 --cfg-split
for v in "$@"; do run gb32_$v --global-batch 32 --variant $v; done
