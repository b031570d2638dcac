"""CPU ORACLE (test infrastructure, NOT product code): decoder half of the first-stage AutoencoderKL.

fp32 functional restatement, driven by a reference-format state_dict (keys `post_quant_conv.*`, `decoder.*`), of
  * AutoencoderKL.decode / encode R/lvdm/models/autoencoder.py:97-106 (encode returns the moments [mean | logvar] of the posterior)
  * Decoder.forward               R/lvdm/modules/networks/ae_modules.py:471-583;  Encoder.forward :364-468, Downsample :90-108
  * ResnetBlock / AttnBlock / Upsample / Normalize   ae_modules.py:151-209, 26-80, 111-126, 16-17 (GroupNorm 32, eps 1e-6; swish)
Pinned by tests/golden/vae_small.npz (output of the unmodified reference classes, oracle/refgen/make_golden.py)."""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F


def _gn(sd, n, x):
    return F.group_norm(x, 32, sd[n + ".weight"], sd[n + ".bias"], 1e-6)


def _conv(sd, n, x, pad):
    return F.conv2d(x, sd[n + ".weight"], sd[n + ".bias"], padding=pad)


def _res(sd, n, x):
    h = _conv(sd, n + ".conv1", F.silu(_gn(sd, n + ".norm1", x)), 1)
    h = _conv(sd, n + ".conv2", F.silu(_gn(sd, n + ".norm2", h)), 1)
    if n + ".nin_shortcut.weight" in sd:
        x = _conv(sd, n + ".nin_shortcut", x, 0)
    return x + h


def _attn(sd, n, x):
    b, c, hh, ww = x.shape
    h = _gn(sd, n + ".norm", x)
    q, k, v = (_conv(sd, f"{n}.{t}", h, 0).reshape(b, c, hh * ww) for t in ("q", "k", "v"))
    w = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** -0.5), dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, n + ".proj_out", h, 0)


def decode(sd: Dict[str, torch.Tensor], z: torch.Tensor, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks: int = 2) -> torch.Tensor:
    """z [N, z_channels, h, w] (N = batch x frames) -> images [N, out_ch, 8h, 8w] for four resolution levels."""
    h = _conv(sd, "post_quant_conv", z, 0)
    h = _conv(sd, "decoder.conv_in", h, 1)
    h = _res(sd, "decoder.mid.block_1", h)
    h = _attn(sd, "decoder.mid.attn_1", h)
    h = _res(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(len(ch_mult))):
        for i in range(num_res_blocks + 1):
            h = _res(sd, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), 1)
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.norm_out", h)), 1)


def encode_moments(sd: Dict[str, torch.Tensor], x: torch.Tensor, ch_mult: Sequence[int] = (1, 2, 4, 4), num_res_blocks: int = 2) -> torch.Tensor:
    """x [N, 3, H, W] -> moments [N, 2*z_channels, H/8, W/8] = quant_conv(encoder(x)) (mean = first half, logvar = second half)."""
    h = _conv(sd, "encoder.conv_in", x, 1)
    for lvl in range(len(ch_mult)):
        for i in range(num_res_blocks):
            h = _res(sd, f"encoder.down.{lvl}.block.{i}", h)
        if lvl != len(ch_mult) - 1:
            n = f"encoder.down.{lvl}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[n + ".weight"], sd[n + ".bias"], stride=2)
    h = _res(sd, "encoder.mid.block_1", h)
    h = _attn(sd, "encoder.mid.attn_1", h)
    h = _res(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.norm_out", h)), 1)
    return _conv(sd, "quant_conv", h, 0)
