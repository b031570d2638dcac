"""Camera pose encoder on the B200 kernels: turns the Pluecker embedding [B, 6, F, H, W] of a sample's
camera trajectory into the four feature maps (320 / 640 / 1280 / 1280 channels at 1/8 .. 1/64 resolution) that the UNet's
encoder blocks add to their hidden states (`pluker_embedding_features`).  Once per sample.

Mirrors R/CamContextI2V/model/modules/camera_pose_encoder.py:295-376 (`CameraPoseEncoder`: constructor keywords, parameter names
and shapes so a reference checkpoint loads with load_state_dict, forward signature and output format), with its `ResnetBlock`
(:236-268), `Downsample` (:212-233), `TemporalTransformerBlock` (:15-78), `PositionalEncoding` (:81-98) and
`TemporalSelfAttention` (:101-160).  The two classes that file imports from `diffusers` (`Attention`, `FeedForward`; the
library is neither vendored in the reference nor installed here) are restated from the published library - parameter names
`to_q / to_k / to_v / to_out.0`, `ff.net.0.proj / ff.net.2` - see oracle/pose_encoder_oracle.py for what that means for parity.

Kernels: `c2v_pixel_unshuffle_cl`, `c2v_avgpool2_cl`, `c2v_relu`, `c2v_attention_temporal_hd` (head dims 40 / 80 / 160) and the
hot path's own `c2v_gemm` (3x3 implicit-GEMM conv, 1x1 conv, fused q|k|v projection, GEGLU epilogue, residual epilogue),
`c2v_layernorm` (its `add` output carries the temporal position encoding), `c2v_cast_bf16`, `c2v_from_channels_last`.
Activations are channel-last rows ordered (b, f, y, x) throughout, so neither of the reference's two rearranges per block
('(b f) c h w -> (b h w) f c' and back) moves any data: the temporal attention kernel strides over f in place.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .modules import _Prepared, _bf16, _conv3x3_pack, _f32

F32 = torch.float32


class PositionalEncoding(nn.Module):
    """camera_pose_encoder.py:81-98: buffer `pe` [1, max_len, d_model]."""

    def __init__(self, d_model, dropout=0.0, max_len=32):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class TemporalSelfAttention(nn.Module):
    """Parameter holder of camera_pose_encoder.py:101-160 over diffusers' Attention (self-attention, no q/k/v bias, out bias)."""

    def __init__(self, query_dim, heads=8, dim_head=64, bias=False, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 rescale_output_factor=1.0):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.rescale_output_factor = heads, dim_head, rescale_output_factor
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.pos_encoder = PositionalEncoding(query_dim, max_len=temporal_position_encoding_max_len) if temporal_position_encoding else None


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """Parameter holder of diffusers' FeedForward(dim, activation_fn="geglu"): net.0.proj, net.2."""

    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class TemporalTransformerBlock(_Prepared):
    """camera_pose_encoder.py:15-78."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0,
                 cross_attention_dim=None, attention_bias=False, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 rescale_output_factor=1.0, **unused):
        super().__init__()
        if any(t != "Temporal_Self" for t in attention_block_types):
            raise NotImplementedError(f"attention_block_types {attention_block_types}: the reference class asserts Temporal_Self too")
        if num_attention_heads * attention_head_dim != dim:
            raise NotImplementedError("inner dim != dim")
        self.dim, self.heads, self.head_dim = dim, num_attention_heads, attention_head_dim
        self.attention_blocks = nn.ModuleList([
            TemporalSelfAttention(dim, num_attention_heads, attention_head_dim, attention_bias, temporal_position_encoding,
                                  temporal_position_encoding_max_len, rescale_output_factor) for _ in attention_block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in attention_block_types])
        self.ff = FeedForward(dim)
        self.ff_norm = nn.LayerNorm(dim)
        self._pe_rows = {}

    def _prepare(self):
        p = {"attn": []}
        for blk, norm in zip(self.attention_blocks, self.norms):
            s = 1.0 / blk.rescale_output_factor
            qkv_b = None
            if blk.to_q.bias is not None:
                qkv_b = _f32(torch.cat([blk.to_q.bias, blk.to_k.bias, blk.to_v.bias]))
            p["attn"].append({"g": _f32(norm.weight), "b": _f32(norm.bias),
                              "wqkv": _bf16(torch.cat([blk.to_q.weight, blk.to_k.weight, blk.to_v.weight], dim=0)), "bqkv": qkv_b,
                              "wo": _bf16(blk.to_out[0].weight.detach() * s), "bo": _f32(blk.to_out[0].bias.detach() * s),
                              "pe": _f32(blk.pos_encoder.pe[0]) if blk.pos_encoder is not None else None})
        w1, b1 = ops.geglu_interleave(self.ff.net[0].proj.weight.detach(), self.ff.net[0].proj.bias.detach())
        p.update(g=_f32(self.ff_norm.weight), b=_f32(self.ff_norm.bias), w1=_bf16(w1), b1=_f32(b1), w2=_bf16(self.ff.net[2].weight),
                 b2=_f32(self.ff.net[2].bias))
        self._pe_rows = {}
        return p

    def _pe(self, i, pe, B, T, HW):
        """Position encoding broadcast to the row order (b, f, pixel): the `add` operand of the layer-norm kernel."""
        key = (i, B, T, HW)
        if key not in self._pe_rows:
            self._pe_rows[key] = pe[:T].repeat_interleave(HW, dim=0).repeat(B, 1).contiguous()
        return self._pe_rows[key]

    def run(self, x: torch.Tensor, B: int, T: int, HW: int) -> torch.Tensor:
        """x fp32 rows [(b, f, pixel), dim] -> the same; attention over f per (b, pixel)."""
        p = self.pk()
        for i, a in enumerate(p["attn"]):
            if a["pe"] is not None:
                n = ops.layernorm(x, a["g"], a["b"], add=self._pe(i, a["pe"], B, T, HW))[1]
            else:
                n = ops.layernorm(x, a["g"], a["b"])
            qkv = ops.linear(n, a["wqkv"], bias=a["bqkv"], out_dtype=ops.BF16)
            o = ops.attention_temporal_hd(qkv, B, T, HW, self.heads, self.head_dim)
            x = ops.linear(o, a["wo"], bias=a["bo"], residual=x)
        h = ops.geglu_linear(ops.layernorm(x, p["g"], p["b"]), p["w1"], p["b1"])
        return ops.linear(h, p["w2"], bias=p["b2"], residual=x)


class Downsample(nn.Module):
    """camera_pose_encoder.py:212-233, dims=2."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if use_conv or dims != 2:
            raise NotImplementedError("Downsample(use_conv=True): not used by any shipped pose_encoder_config")
        self.channels = channels


class ResnetBlock(_Prepared):
    """camera_pose_encoder.py:236-268."""

    def __init__(self, in_c, out_c, down, ksize=3, sk=False, use_conv=True):
        super().__init__()
        if ksize not in (1, 3):
            raise NotImplementedError(f"ksize {ksize}")
        ps = ksize // 2
        self.in_conv = nn.Conv2d(in_c, out_c, ksize, 1, ps) if (in_c != out_c or not sk) else None
        self.block1 = nn.Conv2d(out_c, out_c, 3, 1, 1)
        self.block2 = nn.Conv2d(out_c, out_c, ksize, 1, ps)
        self.skep = nn.Conv2d(in_c, out_c, ksize, 1, ps) if not sk else None
        self.down = down
        if down:
            self.down_opt = Downsample(in_c, use_conv=use_conv)

    @staticmethod
    def _pack(conv):
        if conv is None:
            return None
        if conv.kernel_size[0] == 3:
            return _conv3x3_pack(conv) + (3,)
        return _bf16(conv.weight.detach().flatten(1)), _f32(conv.bias), 1

    def _prepare(self):
        return {k: self._pack(getattr(self, k)) for k in ("in_conv", "block1", "block2", "skep")}

    @staticmethod
    def _conv(pk, a, N, H, W, residual=None, out_dtype=F32):
        w, b, k = pk
        if k == 3:
            return ops.conv3x3(a, w, N, H, W, bias=b, residual=residual, out_dtype=out_dtype)
        return ops.linear(a, w, bias=b, residual=residual, out_dtype=out_dtype)

    def run(self, x: torch.Tensor, x16, N: int, H: int, W: int):
        """x fp32 rows [(n, y, x), in_c] (x16: its 16-bit copy or None) -> (fp32 rows [(n, y', x'), out_c], y', x')."""
        p = self.pk()
        if self.down:
            x, x16 = ops.avgpool2_cl(x, N, H, W)
            H, W = H // 2, W // 2
        elif x16 is None:
            x16 = ops.cast_bf16(x)
        if p["in_conv"] is not None:
            x = self._conv(p["in_conv"], x16, N, H, W)
            x16 = ops.cast_bf16(x)
        h = ops.relu_(self._conv(p["block1"], x16, N, H, W, out_dtype=ops.BF16))
        res = self._conv(p["skep"], x16, N, H, W) if p["skep"] is not None else x
        return self._conv(p["block2"], h, N, H, W, residual=res), H, W


class CameraPoseEncoder(_Prepared):
    def __init__(self, downscale_factor, channels=[320, 640, 1280, 1280], nums_rb=3, cin=64, ksize=3, sk=False, use_conv=True,
                 compression_factor=1, temporal_attention_nhead=8, attention_block_types=("Temporal_Self",), temporal_position_encoding=False,
                 temporal_position_encoding_max_len=16, rescale_output_factor=1.0):
        super().__init__()
        self.downscale_factor, self.channels, self.nums_rb = downscale_factor, list(channels), nums_rb
        self.encoder_down_conv_blocks = nn.ModuleList()
        self.encoder_down_attention_blocks = nn.ModuleList()
        for i in range(len(channels)):
            convs, attns = nn.ModuleList(), nn.ModuleList()
            mid = int(channels[i] / compression_factor)
            for j in range(nums_rb):                                                      # camera_pose_encoder.py:319-334
                if j == 0:
                    in_dim, out_dim, down = (channels[i - 1] if i != 0 else channels[0]), mid, i != 0
                elif j == nums_rb - 1:
                    in_dim, out_dim, down = mid, channels[i], False
                else:
                    in_dim, out_dim, down = mid, mid, False
                convs.append(ResnetBlock(in_dim, out_dim, down=down, ksize=ksize, sk=sk, use_conv=use_conv))
                attns.append(TemporalTransformerBlock(dim=out_dim, num_attention_heads=temporal_attention_nhead,
                                                      attention_head_dim=int(out_dim / temporal_attention_nhead),
                                                      attention_block_types=tuple(attention_block_types),
                                                      temporal_position_encoding=temporal_position_encoding,
                                                      temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                                      rescale_output_factor=rescale_output_factor))
            self.encoder_down_conv_blocks.append(convs)
            self.encoder_down_attention_blocks.append(attns)
        self.encoder_conv_in = nn.Conv2d(cin, channels[0], 3, 1, 1)

    @property
    def dtype(self) -> torch.dtype:
        return self.encoder_conv_in.weight.dtype

    def _prepare(self):
        w, b = _conv3x3_pack(self.encoder_conv_in)
        return {"w_in": w, "b_in": b}

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """x fp32 [B, 6, F, H, W] -> list of fp32 [(B F), C_i, H_i, W_i] (the reference's 'bf c h w' format)."""
        p = self.pk()
        B, _, T, H, W = x.shape
        r = self.downscale_factor
        h, w, N = H // r, W // r, B * T
        x16 = ops.pixel_unshuffle_cl(_f32(x), r)
        xs = ops.conv3x3(x16, p["w_in"], N, h, w, bias=p["b_in"])
        x16 = None
        feats = []
        for convs, attns in zip(self.encoder_down_conv_blocks, self.encoder_down_attention_blocks):
            for res, att in zip(convs, attns):
                xs, h, w = res.run(xs, x16, N, h, w)
                xs = att.run(xs, B, T, h * w)
                x16 = None
            feats.append(ops.from_channels_last(xs, N, xs.shape[1], h * w).view(N, xs.shape[1], h, w))
        return feats
