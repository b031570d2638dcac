"""CPU ORACLE (test infrastructure, NOT product code): MultiLatentEpipolarAdaptor and its conditional epipolar mask.

fp32 functional restatement, driven by a reference-format state_dict, of
  * MultiLatentEpipolarAdaptor._forward        R/model/modules/adaptors.py:137-182   (ctor :36-131)
  * EpipolarCrossAttention.efficient_forward   R/model/modules/epipolar.py:75-102    (register tokens prepended BEFORE to_k / to_v)
  * resampler FeedForward                      R/lvdm/modules/encoders/resampler.py:31-38  (LayerNorm, Linear, GELU(erf), Linear; no bias)
  * compute_conditional_epipolar_mask          R/model/camcontexti2v.py:493-521  (+ get_pairwise_relative_pose, base.py:200-217)
Pinned by tests/golden/adaptor_small.npz (outputs of the unmodified reference classes, oracle/refgen/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import epipolar_mask
from .unet_oracle import softmax_attention, timestep_embedding


def conditional_fundamental_matrices(K: torch.Tensor, w2c: torch.Tensor, w2c_cond: torch.Tensor, cond_frame_index: Optional[torch.Tensor]):
    """K [B,T,3,3], w2c [B,T,4,4] (target frames), w2c_cond [B,n,4,4] (additional context frames) -> F [B, T, C, 3, 3] with
    C = 1 + n (the reference frame first) - camcontexti2v.py:495-516.  No perturbation is applied on this path."""
    c2w = w2c.float().inverse()
    c2w_cond = w2c_cond.float().inverse()
    if cond_frame_index is not None:
        c2w_cond = torch.cat((c2w[torch.arange(len(cond_frame_index)), cond_frame_index].unsqueeze(1), c2w_cond), dim=1)
    rel = c2w_cond.inverse()[:, :, None] @ c2w[:, None]            # [B, C, T]: inv(c2w_cond[c]) @ c2w[t]   (mode='left')
    rel = rel.transpose(1, 2)                                      # [B, T, C]
    R, t = rel[..., :3, :3], rel[..., :3, 3:4]
    Kb = K.float()[:, :, None]                                     # the TARGET frame's intrinsics for every context frame
    E = torch.cross(t, R, dim=-2)
    Kinv = torch.inverse(Kb)
    return Kinv.transpose(-1, -2) @ E @ Kinv


def conditional_epipolar_mask(K, w2c, w2c_cond, cond_frame_index, H: int, W: int, downsample: int = 8) -> torch.Tensor:
    """bool [B, T*h*w, C*h*w], h = H // downsample."""
    Fm = conditional_fundamental_matrices(K, w2c, w2c_cond, cond_frame_index)
    return epipolar_mask(Fm, H // downsample, W // downsample, downsample)


def adaptor_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, mask: Optional[torch.Tensor], depth: int, video_length: int = 16,
                    heads: int = 8, timestep_embedding_dim: int = 32, timestep_embedding_type: str = "sinusoidal_embedded") -> torch.Tensor:
    """x [B, C*h*w, embedding_dim] (latents of the reference + context frames as tokens), mask bool [B, T*h*w, C*h*w] ->
    [B, T*h*w, output_dim]."""
    B = x.shape[0]
    latents = sd["latents"].repeat(B, 1, 1)
    D = latents.shape[-1]
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    for i in range(depth):
        a, f = f"layers.{i}.0.", f"layers.{i}.1."
        reg = sd[a + "register_tokens"]
        q = F.linear(latents, sd[a + "to_q.weight"])
        ctx = torch.cat([reg.repeat(B, 1, 1), x], dim=1)
        k, v = F.linear(ctx, sd[a + "to_k.weight"]), F.linear(ctx, sd[a + "to_v.weight"])
        m = F.pad(mask, (reg.shape[1], 0), value=True) if mask is not None else None
        out = softmax_attention(q, k, v, heads, m)
        latents = F.linear(out, sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"]) + latents
        h = F.layer_norm(latents, (D,), sd[f + "0.weight"], sd[f + "0.bias"], 1e-5)
        h = F.linear(F.gelu(F.linear(h, sd[f + "1.weight"])), sd[f + "3.weight"])
        latents = h + latents
    if timestep_embedding_type != "none":
        t_emb = timestep_embedding(torch.arange(video_length), timestep_embedding_dim if timestep_embedding_type.endswith("embedded") else D)
        if timestep_embedding_type == "sinusoidal_embedded":
            t_emb = F.linear(F.silu(F.linear(t_emb, sd["timestep_embedding_func.0.weight"], sd["timestep_embedding_func.0.bias"])),
                             sd["timestep_embedding_func.2.weight"], sd["timestep_embedding_func.2.bias"])
        L = latents.shape[1] // video_length
        latents = latents + t_emb[None, :, None, :].expand(B, video_length, L, D).reshape(B, video_length * L, D)
    y = F.linear(latents, sd["proj_out.weight"], sd["proj_out.bias"])
    return F.layer_norm(y, (y.shape[-1],), sd["norm_out.weight"], sd["norm_out.bias"], 1e-5)
