"""2-GPU check of the CFG-split mode (parallel.CfgPair): the pair's step must equal the single-GPU step bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/cfg_split_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402
from camc2v_b200.parallel import make_cfg_pairs  # noqa: E402
from camc2v_b200.sampler import DDIMSampler  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    pair = make_cfg_pairs(rank, dist.get_world_size())
    cfg = UNetConfig()
    model, sampler, host, cam_host, _ = bench.build_workload(cfg, 1, device)
    cond, uc, static, _ = bench.to_device_conditioning(host, cam_host, device)
    kw = dict(unconditional_guidance_scale=3.5, unconditional_conditioning=uc, guidance_rescale=0.7, fs=static["fs"],
              enable_camera_condition=True)
    x = host["x"].to(device)
    ts = torch.full((1,), 999, device=device, dtype=torch.long)
    noise = pair.noise(x.shape, device)
    full, _ = sampler.p_sample_ddim(x, cond, ts, index=24, noise=noise, use_cuda_graph=True, **kw)
    split_sampler = DDIMSampler(model, cfg_pair=pair)
    split_sampler.make_schedule(25, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    for graph in (False, True):
        got, _ = split_sampler.p_sample_ddim(x, cond, ts, index=24, noise=noise, use_cuda_graph=graph, **kw)
        assert torch.equal(got, full), f"rank {rank} graph={graph}: max diff {(got - full).abs().max().item()}"
    # without explicit noise both ranks draw the same eta-noise from the pair generator
    a, _ = split_sampler.p_sample_ddim(x, cond, ts, index=24, use_cuda_graph=True, **kw)
    both = [torch.empty_like(a) for _ in range(2)]
    dist.all_gather(both, a)
    assert torch.equal(both[0], both[1])
    print(f"rank {rank} (role {pair.role}): CFG-split step == single-GPU step, bit-exact; pair latents identical", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
