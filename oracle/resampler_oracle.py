"""CPU ORACLE (test infrastructure, NOT product code): the Resampler image-token projector.

fp32 functional restatement of R/lvdm/modules/encoders/resampler.py:52-166 (PerceiverAttention :52-97, FeedForward :31-38,
Resampler :100-166), driven by a reference-format state_dict.  Pinned by tests/golden/resampler_small.npz (output of the
unmodified reference class, oracle/refgen/make_golden.py)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .unet_oracle import softmax_attention, timestep_embedding


def resampler_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, depth: int, heads: int, video_length: int = 16,
                      use_timestep_emb: bool = True) -> torch.Tensor:
    """x [B, n, embedding_dim] -> [B, num_queries * video_length, output_dim]."""
    B = x.shape[0]
    latents = sd["latents"].repeat(B, 1, 1)
    D = latents.shape[-1]
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    for i in range(depth):
        a, f = f"layers.{i}.0.", f"layers.{i}.1."
        xn = F.layer_norm(x, (D,), sd[a + "norm1.weight"], sd[a + "norm1.bias"], 1e-5)
        ln = F.layer_norm(latents, (D,), sd[a + "norm2.weight"], sd[a + "norm2.bias"], 1e-5)
        q = F.linear(ln, sd[a + "to_q.weight"])
        k, v = F.linear(torch.cat((xn, ln), dim=-2), sd[a + "to_kv.weight"]).chunk(2, dim=-1)
        out = softmax_attention(q, k, v, heads)              # (q d^-1/4)(k d^-1/4)^T = q k^T / sqrt(d)
        latents = F.linear(out, sd[a + "to_out.weight"]) + latents
        h = F.layer_norm(latents, (D,), sd[f + "0.weight"], sd[f + "0.bias"], 1e-5)
        latents = F.linear(F.gelu(F.linear(h, sd[f + "1.weight"])), sd[f + "3.weight"]) + latents
    if use_timestep_emb:
        t_emb = timestep_embedding(torch.arange(video_length), D)
        t_emb = F.linear(F.silu(F.linear(t_emb, sd["timestep_embedding_func.0.weight"], sd["timestep_embedding_func.0.bias"])),
                         sd["timestep_embedding_func.2.weight"], sd["timestep_embedding_func.2.bias"])
        L = latents.shape[1] // video_length
        latents = latents + t_emb[None, :, None, :].expand(B, video_length, L, D).reshape(B, video_length * L, D)
    y = F.linear(latents, sd["proj_out.weight"], sd["proj_out.bias"])
    return F.layer_norm(y, (y.shape[-1],), sd["norm_out.weight"], sd["norm_out.bias"], 1e-5)
