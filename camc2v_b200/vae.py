"""Decoder half of the first-stage AutoencoderKL on the B200 kernels: `decode_first_stage`, the step
right after the sampling loop - a conv / GroupNorm / swish network that takes the 16 latent frames [4, 32, 32] of a video to
16 RGB frames [3, 256, 256] (0.75 TFLOP per frame, 12 TFLOP per video: about one DDIM step).

Mirrors R/lvdm/modules/networks/ae_modules.py:471-583 (`Decoder`, with `ResnetBlock` :151-209, `AttnBlock` :26-80, `Upsample`
:111-126) and `AutoencoderKL.decode` (R/lvdm/models/autoencoder.py:103-106): same constructor keywords of `ddconfig`, same
parameter names and shapes for `post_quant_conv.*` and `decoder.*` (the encoder half is not part of this row).  Activations are
channels-last fp32 between blocks, 16-bit operands into every GEMM; all 3x3 convolutions are the implicit-GEMM tcgen05 kernel of
the UNet (4-D TMA maps, now also for images wider than one 128-pixel tile), GroupNorm + swish is the fused norm kernel, the
nearest-2x upsample writes the 16-bit operand of its conv directly, and the single-head 512-wide attention block goes through
two GEMMs (q k^T and P v^T-as-weights) around `c2v_softmax_rows`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .modules import _Prepared, _bf16, _conv3x3_pack, _f32

F32 = torch.float32


def Normalize(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        assert temb_channels == 0 and not conv_shortcut
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1, self.conv1 = Normalize(in_channels), nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2, self.conv2 = Normalize(out_channels), nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def pack(self):
        w1, b1 = _conv3x3_pack(self.conv1)
        w2, b2 = _conv3x3_pack(self.conv2)
        p = dict(g1=_f32(self.norm1.weight), be1=_f32(self.norm1.bias), w1=w1, b1=b1, g2=_f32(self.norm2.weight), be2=_f32(self.norm2.bias),
                 w2=w2, b2=b2)
        if self.in_channels != self.out_channels:
            # 16-bit copy of the unnormalised stream is stored at ops.RESIDUAL_PRESCALE, the weights carry the inverse (fp16 range)
            p["ws"] = _bf16(self.nin_shortcut.weight.reshape(self.out_channels, self.in_channels) * (1.0 / ops.RESIDUAL_PRESCALE))
            p["bs"] = _f32(self.nin_shortcut.bias)
        return p

    @staticmethod
    def run(p, h, N, H, W):
        n = ops.groupnorm(h, p["g1"], p["be1"], N, H * W, 1e-6, True)
        h1 = ops.conv3x3(n, p["w1"], N, H, W, bias=p["b1"])
        n = ops.groupnorm(h1, p["g2"], p["be2"], N, H * W, 1e-6, True)
        skip = ops.linear(ops.cast_bf16(h, ops.RESIDUAL_PRESCALE), p["ws"], bias=p["bs"]) if "ws" in p else h
        return ops.conv3x3(n, p["w2"], N, H, W, bias=p["b2"], residual=skip)


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(in_channels, in_channels, 1, 1, 0) for _ in range(4))

    def pack(self):
        c = self.in_channels
        return dict(g=_f32(self.norm.weight), be=_f32(self.norm.bias),
                    wq=_bf16(self.q.weight.reshape(c, c)), bq=_f32(self.q.bias), wk=_bf16(self.k.weight.reshape(c, c)), bk=_f32(self.k.bias),
                    wv=_bf16(self.v.weight.reshape(c, c)), bv=_f32(self.v.bias),
                    wo=_bf16(self.proj_out.weight.reshape(c, c)), bo=_f32(self.proj_out.bias))

    @staticmethod
    def run(p, h, N, H, W):
        c, L = h.shape[1], H * W
        n = ops.groupnorm(h, p["g"], p["be"], N, L, 1e-6, False)
        q = ops.linear(n, p["wq"], bias=p["bq"], out_dtype=ops.BF16)                      # [N*L, c]
        k = ops.linear(n, p["wk"], bias=p["bk"], out_dtype=ops.BF16)
        out = torch.empty((N * L, c), device=h.device, dtype=ops.BF16)
        for i in range(N):                                                                # one image at a time: [L, L] scores
            qi, ki, ni = q[i * L:(i + 1) * L], k[i * L:(i + 1) * L], n[i * L:(i + 1) * L]
            s = ops.linear(qi, ki)                                                        # fp32 [L, L] = q k^T
            pm = ops.softmax_rows(s, float(c) ** -0.5)
            vt = ops.linear(p["wv"], ni, out_dtype=ops.BF16)                              # [c, L] = W_v n^T  (v^T without its bias)
            # rows of P sum to 1, so the value bias passes straight through:  P (v + 1 b^T) = P v + b^T
            ops.linear(pm, vt, bias=p["bv"], out=out[i * L:(i + 1) * L], out_dtype=ops.BF16)
        return ops.linear(out, p["wo"], bias=p["bo"], residual=h)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        assert with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Decoder(_Prepared):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True, in_channels,
                 resolution, z_channels, give_pre_end=False, tanh_out=False, use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if list(attn_resolutions) or give_pre_end or tanh_out or use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("decoder options the shipped first_stage_config leaves off (camcontexti2v_256.yaml:74-93)")
        self.ch_mult, self.num_res_blocks, self.out_ch, self.z_channels = tuple(ch_mult), num_res_blocks, out_ch, z_channels
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.up = nn.ModuleList()
        for lvl in reversed(range(len(ch_mult))):
            block = nn.ModuleList()
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
            up = nn.Module()
            up.block, up.attn = block, nn.ModuleList()
            if lvl != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        assert with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class Encoder(nn.Module):
    """Parameter holder of ae_modules.py:364-428."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True, in_channels,
                 resolution, z_channels, double_z=True, use_linear_attn=False, attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if list(attn_resolutions) or use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("encoder options the shipped first_stage_config leaves off")
        self.ch_mult, self.num_res_blocks, self.in_channels = tuple(ch_mult), num_res_blocks, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(len(ch_mult)):
            block = nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[lvl], ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out))
                block_in = block_out
            down = nn.Module()
            down.block, down.attn = block, nn.ModuleList()
            if lvl != len(ch_mult) - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class AutoencoderKLEncoder(_Prepared):
    """`encoder` + `quant_conv` of lvdm.models.autoencoder.AutoencoderKL (autoencoder.py:97-101), state_dict-compatible.
    `encode` returns the posterior moments [mean | logvar]; sampling the DiagonalGaussianDistribution stays with the caller."""

    def __init__(self, ddconfig: dict, embed_dim: int = 4):
        super().__init__()
        assert ddconfig.get("double_z", True)
        self.encoder = Encoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)

    def _prepare(self):
        e = self.encoder
        w_in, b_in = _conv3x3_pack(e.conv_in, pad_cin=64)
        # quant_conv (1x1) follows conv_out (3x3) with nothing in between: W_q (W_out * x + b_out) + b_q is one 3x3 conv (exact fold)
        wq = self.quant_conv.weight.detach().float().reshape(self.quant_conv.out_channels, -1)
        wo = e.conv_out.weight.detach().float()
        w_f = torch.einsum("om,mikl->oikl", wq, wo)
        b_f = wq @ e.conv_out.bias.detach().float() + self.quant_conv.bias.detach().float()
        p = dict(w_in=w_in, b_in=b_in, mid1=e.mid.block_1.pack(), attn=e.mid.attn_1.pack(), mid2=e.mid.block_2.pack(),
                 g_out=_f32(e.norm_out.weight), be_out=_f32(e.norm_out.bias),
                 w_out=_bf16(w_f.permute(0, 2, 3, 1).reshape(w_f.shape[0], -1)), b_out=_f32(b_f), down=[])
        for d in e.down:
            u = dict(blocks=[b.pack() for b in d.block])
            if hasattr(d, "downsample"):
                w = d.downsample.conv.weight.detach()
                u["w_dn"], u["b_dn"] = _bf16(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)), _f32(d.downsample.conv.bias)
            p["down"].append(u)
        return p

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [N, 3, H, W] -> moments fp32 [N, 2*embed_dim, H/8, W/8]."""
        p = self.pk()
        e = self.encoder
        N, c, H, W = x.shape
        xin = ops.to_channels_last(_f32(x), N, c, H * W, Cpad=64, dtype=ops.BF16)
        h = ops.conv3x3(xin, p["w_in"], N, H, W, bias=p["b_in"])
        for lvl in range(len(e.ch_mult)):
            u = p["down"][lvl]
            for bp in u["blocks"]:
                h = ResnetBlock.run(bp, h, N, H, W)
            if "w_dn" in u:
                cols = ops.im2col_s2(h, N, H, W, pad_lo=0)                                   # F.pad(x, (0,1,0,1)) + stride-2 conv
                H, W = H // 2, W // 2
                h = ops.linear(cols, u["w_dn"], bias=u["b_dn"])
        h = ResnetBlock.run(p["mid1"], h, N, H, W)
        h = AttnBlock.run(p["attn"], h, N, H, W)
        h = ResnetBlock.run(p["mid2"], h, N, H, W)
        n = ops.groupnorm(h, p["g_out"], p["be_out"], N, H * W, 1e-6, True)
        y = ops.conv3x3(n, p["w_out"], N, H, W, bias=p["b_out"])                              # conv_out and quant_conv, folded
        return ops.from_channels_last(y, N, y.shape[1], H * W).view(N, -1, H, W)

    forward = encode


class AutoencoderKLDecoder(_Prepared):
    """`post_quant_conv` + `decoder` of lvdm.models.autoencoder.AutoencoderKL (autoencoder.py:103-106), state_dict-compatible."""

    def __init__(self, ddconfig: dict, embed_dim: int = 4):
        super().__init__()
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)

    def _prepare(self):
        d = self.decoder
        zc, e = d.z_channels, self.post_quant_conv.in_channels
        # post_quant_conv (1x1, 4 -> 4) folded into conv_in (3x3, 4 -> C): both linear with zero padding only on the 3x3 -> exact
        # for the weights; the folded bias differs at the 1-pixel border (the reference pads the post_quant OUTPUT, bias included,
        # with zeros), so the two convolutions stay separate: 1x1 as a GEMM with the 4 channels padded to 64
        wq = torch.zeros(64, 64, device=self.post_quant_conv.weight.device)
        wq[:zc, :e] = self.post_quant_conv.weight.detach().reshape(zc, e)
        bq = torch.zeros(64, device=wq.device)
        bq[:zc] = self.post_quant_conv.bias.detach()
        w_in, b_in = _conv3x3_pack(d.conv_in, pad_cin=64)
        wo = d.conv_out.weight.detach()
        wo = torch.cat([wo, wo.new_zeros(4 - d.out_ch % 4 if d.out_ch % 4 else 0, *wo.shape[1:])], 0) if d.out_ch % 4 else wo
        bo = torch.cat([d.conv_out.bias.detach(), d.conv_out.bias.new_zeros(wo.shape[0] - d.out_ch)])
        p = dict(wq=_bf16(wq), bq=_f32(bq), w_in=w_in, b_in=b_in, mid1=d.mid.block_1.pack(), attn=d.mid.attn_1.pack(), mid2=d.mid.block_2.pack(),
                 g_out=_f32(d.norm_out.weight), be_out=_f32(d.norm_out.bias),
                 w_out=_bf16(wo.permute(0, 2, 3, 1).reshape(wo.shape[0], -1)), b_out=_f32(bo), up=[])
        for up in d.up:
            u = dict(blocks=[b.pack() for b in up.block])
            if hasattr(up, "upsample"):
                u["w_up"], u["b_up"] = _conv3x3_pack(up.upsample.conv)
            p["up"].append(u)
        return p

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z fp32 [N, embed_dim, h, w] (N = batch x frames) -> fp32 [N, out_ch, 8h, 8w]."""
        p = self.pk()
        d = self.decoder
        N, e, H, W = z.shape
        zin = ops.to_channels_last(_f32(z), N, e, H * W, Cpad=64, dtype=ops.BF16)        # [N*H*W, 64] 16-bit, channels >= e zero
        h = ops.linear(zin, p["wq"], bias=p["bq"], out_dtype=ops.BF16)                         # post_quant_conv
        h = ops.conv3x3(h, p["w_in"], N, H, W, bias=p["b_in"])
        h = ResnetBlock.run(p["mid1"], h, N, H, W)
        h = AttnBlock.run(p["attn"], h, N, H, W)
        h = ResnetBlock.run(p["mid2"], h, N, H, W)
        for lvl in reversed(range(len(d.ch_mult))):
            u = p["up"][lvl]
            for bp in u["blocks"]:
                h = ResnetBlock.run(bp, h, N, H, W)
            if lvl != 0:
                up = ops.upsample2x(h, N, H, W)                                              # 16-bit [N, 2H, 2W, C]
                H, W = 2 * H, 2 * W
                h = ops.conv3x3(up, u["w_up"], N, H, W, bias=u["b_up"])
        n = ops.groupnorm(h, p["g_out"], p["be_out"], N, H * W, 1e-6, True)
        y = ops.conv3x3(n, p["w_out"], N, H, W, bias=p["b_out"])                              # [N*H*W, 4] (3 real channels)
        return ops.from_channels_last(y, N, y.shape[1], H * W).view(N, -1, H, W)[:, : d.out_ch].contiguous()

    forward = decode
