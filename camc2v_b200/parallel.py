"""Multi-GPU plumbing of the hot path.

The path shards naturally by video: every sample's 25-step trajectory is independent (no cross-sample op in the
UNet; all norm statistics are per sample / frame / token), so rank r takes videos[r::world], runs its own
sampling loop with NO step-time communication, and the only collective is one all_gather of the final latents
([B_loc, 4, 16, 32, 32] fp32 = 256 KB per video) over NCCL / NVLink — exactly the reference's inference data
parallelism (R/02_generate_videos.py:173-178, R/main/callbacks.py:241-245) minus Lightning.  One process per GPU,
launched by torchrun; `gloo` is used for the CPU tests of this file.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_videos(videos: Sequence, rank: int, world: int) -> List:
    return list(videos[rank::world])


def gather_latents(local: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """local [B_loc, ...] of videos[rank::world] -> [n_total, ...] in the original video order (valid on every rank)."""
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        n_r = len(range(r, n_total, world))
        out[r::world] = parts[r][:n_r]
    return out
