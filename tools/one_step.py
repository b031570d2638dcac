"""Exactly one CFG + DDIM step (cond pass, uncond pass, fused update) between cudaProfilerStart/Stop, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/one_step.py

Eager launches (no CUDA graph) so that every kernel is one ncu record; two warm-up steps run before the profiled one.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402

device = torch.device("cuda", 0)
cfg = UNetConfig()
model, sampler, host, cam_host, _ = bench.build_workload(cfg, 1, device)
cond, uc, static, _ = bench.to_device_conditioning(host, cam_host, device)
kw = dict(unconditional_guidance_scale=3.5, unconditional_conditioning=uc, guidance_rescale=0.7, fs=static["fs"],
          enable_camera_condition=True, use_cuda_graph=False)
ts_table = np.flip(sampler.ddim_timesteps).copy()
x = host["x"].to(device)
for i in range(2):
    ts = torch.full((1,), int(ts_table[i]), device=device, dtype=torch.long)
    x = sampler.p_sample_ddim(x, cond, ts, index=24 - i, **kw)[0]
torch.cuda.synchronize()
ts = torch.full((1,), int(ts_table[2]), device=device, dtype=torch.long)
noise = torch.randn(x.shape, device=device)
torch.cuda.synchronize()
torch.cuda.profiler.start()
x = sampler.p_sample_ddim(x, cond, ts, index=22, noise=noise, **kw)[0]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("one step done, finite:", bool(torch.isfinite(x).all()))
