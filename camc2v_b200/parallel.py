"""Multi-GPU plumbing of the hot path.

The path shards naturally by video: every sample's 25-step trajectory is independent (no cross-sample op in the
UNet; all norm statistics are per sample / frame / token), so rank r takes videos[r::world], runs its own
sampling loop with NO step-time communication, and the only collective is one all_gather of the final latents
([B_loc, 4, 16, 32, 32] fp32 = 256 KB per video) over NCCL / NVLink — exactly the reference's inference data
parallelism (R/02_generate_videos.py:173-178, R/main/callbacks.py:241-245) minus Lightning.  One process per GPU,
launched by torchrun; `gloo` is used for the CPU tests of this file.

Second axis (BASELINE config 4, SURVEY.md §8e "CFG halves"): the conditional and the unconditional UNet pass of one CFG
step are independent, so a PAIR of ranks can share one video - role 0 runs the conditional pass, role 1 the
unconditional one, the two noise predictions ([B_loc, 4, 16, 32, 32] fp32, 256 KB per video) are exchanged with one
2-rank all_gather per step over NVLink, and both ranks then apply the same fused CFG + DDIM update with the same noise,
so their latents stay bit-identical without any further traffic.  This halves the latency of a single video (latency
mode: fewer videos than GPUs); for throughput plain video sharding is better and remains the default.
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


class CfgPair:
    """The two ranks (2k, 2k+1) that split the CFG halves of the videos of pair k."""

    def __init__(self, rank: int, world: int, group=None, seed: int = 20230211):
        if world % 2 != 0:
            raise ValueError("CFG-split needs an even number of ranks")
        self.rank, self.world = rank, world
        self.role = rank % 2            # 0: conditional pass, 1: unconditional pass
        self.pair_id = rank // 2
        self.n_pairs = world // 2
        self.group = group
        self.seed = seed + self.pair_id
        self._gen = {}

    def exchange(self, e_local: torch.Tensor):
        """e_local = this rank's noise prediction -> (e_cond, e_uncond) on both ranks (one 2-rank all_gather)."""
        n = e_local.shape[0]
        both = torch.empty((2 * n,) + tuple(e_local.shape[1:]), dtype=e_local.dtype, device=e_local.device)
        dist.all_gather_into_tensor(both, e_local.contiguous(), group=self.group)
        return both[:n], both[n:]

    def noise(self, shape, device) -> torch.Tensor:
        """The eta-noise of ddim.py:340, drawn from a generator both ranks of the pair seed identically."""
        key = str(device)
        g = self._gen.get(key)
        if g is None:
            g = self._gen[key] = torch.Generator(device=device)
            g.manual_seed(self.seed)
        return torch.randn(tuple(shape), generator=g, device=device)


def make_cfg_pairs(rank: int, world: int, seed: int = 20230211) -> CfgPair:
    """Create the pair groups (collectively: every rank creates every group, as torch.distributed requires)."""
    mine = None
    for k in range(world // 2):
        g = dist.new_group([2 * k, 2 * k + 1])
        if k == rank // 2:
            mine = g
    return CfgPair(rank, world, mine, seed)


def shard_videos_cfg_split(videos: Sequence, rank: int, world: int) -> List:
    """Videos of this rank's pair (both ranks of a pair get the same list)."""
    return list(videos[rank // 2::world // 2])
