"""Seeded synthetic inputs shared by the golden generator (oracle/refgen/make_golden.py), the tests and bench.py."""
from __future__ import annotations

import torch

from . import synth
from .config import UNetConfig


def synth_unet_inputs(cfg: UNetConfig, hw: int, n_ctx_frames: int, tag: str, B: int = 1, seed: int = 7) -> dict:
    """Inputs of one UNet pass: latent, channel-concat condition, cond / uncond cross-attention
    context (77 text + 256 image tokens per reference/context frame), per-level Pluecker features, fs."""
    T = cfg.temporal_length
    mc = cfg.model_channels
    x = synth.synth_tensor(f"{tag}.x", (B, 4, T, hw, hw), seed)
    c_concat = synth.synth_tensor(f"{tag}.c_concat", (B, 4, T, hw, hw), seed)
    ctx_cond = synth.synth_tensor(f"{tag}.ctx_cond", (B, 77 + 256 * (1 + n_ctx_frames), cfg.context_dim), seed)
    ctx_uncond = synth.synth_tensor(f"{tag}.ctx_uncond", (B, 77 + 256, cfg.context_dim), seed)
    chans = [mc * m for m in cfg.channel_mult]
    pf = [synth.synth_tensor(f"{tag}.pluker{i}", (B, c, T, hw >> i, hw >> i), seed, std=0.1) for i, c in enumerate(chans)]
    return dict(x=x, c_concat=c_concat, ctx_cond=ctx_cond, ctx_uncond=ctx_uncond, pluker=pf, fs=torch.full((B,), 3, dtype=torch.long))
