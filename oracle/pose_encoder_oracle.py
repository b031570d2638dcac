"""CPU oracle (TEST INFRASTRUCTURE ONLY - never imported by the product path) of the camera pose encoder, SURVEY f-2.

Plain fp32 torch restatement of CamContextI2V/model/modules/camera_pose_encoder.py:295-376 (`CameraPoseEncoder.forward`),
its `ResnetBlock` (:236-268), `Downsample` (:212-233), `TemporalTransformerBlock` (:15-78), `PositionalEncoding` (:81-98) and
`TemporalSelfAttention` (:101-160) on a reference-format state dict.

PARITY PARTLY UNPINNED.  Two classes the reference file builds on come from a third-party dependency that is absent from
/root/reference and not installed in this image: `diffusers.models.attention_processor.Attention` and
`diffusers.models.attention.FeedForward` (requirements.txt: `diffusers`, unpinned).  They are restated here from the published
library (Attention with the default AttnProcessor2_0: to_q / to_k / to_v without bias, heads split, softmax(q k^T / sqrt(dim_head)) v,
to_out[0] Linear with bias, to_out[1] Dropout, / rescale_output_factor; FeedForward(dim, activation_fn="geglu"): net[0] = GEGLU with
proj = Linear(dim, 8 dim), hidden * gelu(gate); net[1] Dropout; net[2] = Linear(4 dim, dim)).  Everything else in the file is the
reference's own code and IS pinned: tests/golden/pose_encoder_small.npz is produced by the reference's CameraPoseEncoder itself, run
with a stub `diffusers` module that holds these two restated classes (oracle/refgen/diffusers_stub.py), so structure, parameter
names, rearranges, positional encoding, resnet blocks and pooling are the reference's.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def positional_encoding(d_model: int, max_len: int) -> torch.Tensor:
    """camera_pose_encoder.py:88-94 -> [max_len, d_model]."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _attention(sd, pre, x, heads):
    """diffusers Attention (self-attention, AttnProcessor2_0) with the TemporalSelfAttention position encoding (:133-134)."""
    if pre + "pos_encoder.pe" in sd:
        x = x + sd[pre + "pos_encoder.pe"][:, :x.shape[1]]
    B, L, C = x.shape
    q, k, v = (F.linear(x, sd[pre + n + ".weight"]).view(B, L, heads, C // heads).transpose(1, 2) for n in ("to_q", "to_k", "to_v"))
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(C // heads), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, L, C)
    return F.linear(a, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def _transformer(sd, pre, x, heads, n_attn):
    """TemporalTransformerBlock.forward (:64-78); x [(b h w), f, c]."""
    C = x.shape[-1]
    for i in range(n_attn):
        n = F.layer_norm(x, (C,), sd[f"{pre}norms.{i}.weight"], sd[f"{pre}norms.{i}.bias"])
        x = _attention(sd, f"{pre}attention_blocks.{i}.", n, heads) + x
    n = F.layer_norm(x, (C,), sd[pre + "ff_norm.weight"], sd[pre + "ff_norm.bias"])
    h, g = F.linear(n, sd[pre + "ff.net.0.proj.weight"], sd[pre + "ff.net.0.proj.bias"]).chunk(2, dim=-1)
    return F.linear(h * F.gelu(g), sd[pre + "ff.net.2.weight"], sd[pre + "ff.net.2.bias"]) + x


def _resnet(sd, pre, x, down):
    """ResnetBlock.forward (:254-268) with whatever of in_conv / skep the state dict holds."""
    if down:
        if pre + "down_opt.op.weight" in sd:
            x = F.conv2d(x, sd[pre + "down_opt.op.weight"], sd[pre + "down_opt.op.bias"], stride=2, padding=1)
        else:
            x = F.avg_pool2d(x, 2, 2)
    if pre + "in_conv.weight" in sd:
        w = sd[pre + "in_conv.weight"]
        x = F.conv2d(x, w, sd[pre + "in_conv.bias"], padding=w.shape[-1] // 2)
    h = F.relu(F.conv2d(x, sd[pre + "block1.weight"], sd[pre + "block1.bias"], padding=1))
    w = sd[pre + "block2.weight"]
    h = F.conv2d(h, w, sd[pre + "block2.bias"], padding=w.shape[-1] // 2)
    if pre + "skep.weight" in sd:
        w = sd[pre + "skep.weight"]
        return h + F.conv2d(x, w, sd[pre + "skep.bias"], padding=w.shape[-1] // 2)
    return h + x


def pose_encoder_forward(sd, x, downscale_factor=8, n_levels=4, nums_rb=2, heads=8, n_attn=1):
    """CameraPoseEncoder.forward (:357-376): x [B, 6, F, H, W] -> list of [(B F), C_i, H/(8 2^i), W/(8 2^i)]."""
    sd = {k: v.float() for k, v in sd.items()}
    B, _, T, H, W = x.shape
    x = x.float().permute(0, 2, 1, 3, 4).reshape(B * T, -1, H, W)
    x = F.pixel_unshuffle(x, downscale_factor)
    x = F.conv2d(x, sd["encoder_conv_in.weight"], sd["encoder_conv_in.bias"], padding=1)
    feats = []
    for i in range(n_levels):
        for j in range(nums_rb):
            x = _resnet(sd, f"encoder_down_conv_blocks.{i}.{j}.", x, down=(j == 0 and i != 0))
            C, h, w = x.shape[1:]
            s = x.view(B, T, C, h * w).permute(0, 3, 1, 2).reshape(B * h * w, T, C)
            s = _transformer(sd, f"encoder_down_attention_blocks.{i}.{j}.", s, heads, n_attn)
            x = s.view(B, h * w, T, C).permute(0, 2, 3, 1).reshape(B * T, C, h, w)
        feats.append(x)
    return feats
