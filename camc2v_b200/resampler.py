"""Resampler image-token projector on the B200 kernels: 4 perceiver layers that turn the 257 CLIP image
tokens of a frame into 16 x 16 = 256 context tokens of width 1024 (the `77 + 256 (1 + n)` image part of the UNet's
cross-attention context).  Once per reference / context frame.

Mirrors R/lvdm/modules/encoders/resampler.py:100-166 (constructor keywords, parameter names and shapes, forward signature).
Every op runs on the hot path's kernels: `c2v_layernorm`, `c2v_gemm` (GELU epilogue for the FeedForward; the per-frame
embedding folded into a row bias of the proj_out GEMM), `c2v_attention` (dense, ragged 513-key rows), `c2v_copy_rows`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .adaptor import FeedForward
from .modules import _Prepared, _bf16, _f32, _pad_cols

F32 = torch.float32


class PerceiverAttention(nn.Module):
    """Parameter holder of resampler.py:52-66."""

    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        assert dim_head == 64, "the attention kernel is specialised for head dim 64"
        inner = dim_head * heads
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


class Resampler(_Prepared):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024, ff_mult=4,
                 video_length=None, use_timestep_emb=False):
        super().__init__()
        self.num_queries, self.video_length, self.use_timestep_emb, self.dim = num_queries, video_length, use_timestep_emb, dim
        n_lat = num_queries * video_length if video_length is not None else num_queries
        self.latents = nn.Parameter(torch.randn(1, n_lat, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList([nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads), FeedForward(dim=dim, mult=ff_mult)])
                                     for _ in range(depth)])
        if use_timestep_emb:
            assert video_length is not None
            self.timestep_embedding_func = nn.Sequential(nn.Linear(dim, dim), nn.SiLU(), nn.Linear(dim, dim))

    def _prepare(self):
        dev = self.latents.device
        p = {"w_in": _bf16(_pad_cols(self.proj_in.weight.detach(), 64)), "b_in": _f32(self.proj_in.bias), "w_out": _bf16(self.proj_out.weight),
             "b_out": _f32(self.proj_out.bias), "g_out": _f32(self.norm_out.weight), "be_out": _f32(self.norm_out.bias), "layers": []}
        for attn, ff in self.layers:
            p["layers"].append({"g1": _f32(attn.norm1.weight), "b1": _f32(attn.norm1.bias), "g2": _f32(attn.norm2.weight), "b2": _f32(attn.norm2.bias),
                                "wq": _bf16(attn.to_q.weight), "wkv": _bf16(attn.to_kv.weight), "wo": _bf16(attn.to_out.weight),
                                "g": _f32(ff[0].weight), "b": _f32(ff[0].bias), "w1": _bf16(ff[1].weight), "w2": _bf16(ff[3].weight)})
        if self.use_timestep_emb:
            # frame embedding (a constant [T, dim]) folded into a per-frame row bias of proj_out: proj_out(l + e_t) = proj_out(l) + W e_t
            t = torch.arange(self.video_length, device=dev, dtype=torch.long)
            f0, f2 = self.timestep_embedding_func[0], self.timestep_embedding_func[2]
            e = ops.skinny_linear(ops.timestep_embedding(t, self.dim), _bf16(f0.weight), _f32(f0.bias), False)
            e = ops.skinny_linear(e, _bf16(f2.weight), _f32(f2.bias), True)
            p["rowbias"] = ops.skinny_linear(e, _bf16(self.proj_out.weight), None, False).contiguous()
        return p

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, n, embedding_dim] fp32 -> [B, num_queries * video_length, output_dim] fp32."""
        p = self.pk()
        B, n, E = x.shape
        Lq, D = self.latents.shape[1], self.dim
        heads = self.layers[0][0].heads
        inner = heads * 64
        Ep = p["w_in"].shape[1]
        xin = torch.zeros((B * n, Ep), device=x.device, dtype=ops.BF16)
        xin[:, :E] = _f32(x).view(B * n, E)
        xs = ops.linear(xin, p["w_in"], bias=p["b_in"])                                       # fp32 [B*n, D], constant over the layers
        lat = _f32(self.latents).expand(B, Lq, D).reshape(B * Lq, D).contiguous()             # fp32 residual stream
        kv_in = torch.empty((B, n + Lq, D), device=x.device, dtype=ops.BF16)                 # cat((norm1(x), norm2(latents)), dim=-2)
        for lp in p["layers"]:
            xn = ops.layernorm(xs, lp["g1"], lp["b1"])
            ln = ops.layernorm(lat, lp["g2"], lp["b2"])
            for b in range(B):
                ops.copy_rows(xn[b * n:(b + 1) * n], kv_in[b, :n])
                ops.copy_rows(ln[b * Lq:(b + 1) * Lq], kv_in[b, n:])
            q = ops.linear(ln, lp["wq"], out_dtype=ops.BF16)
            kv = ops.linear(kv_in.view(B * (n + Lq), D), lp["wkv"], out_dtype=ops.BF16)      # [B*(n+Lq), 2*inner]
            a = ops.attention(q, kv[:, :inner], kv[:, inner:], B, Lq, n + Lq, heads)
            lat = ops.linear(a, lp["wo"], residual=lat)
            h = ops.linear(ops.layernorm(lat, lp["g"], lp["b"]), lp["w1"], out_dtype=ops.BF16, gelu=True)
            lat = ops.linear(h, lp["w2"], residual=lat)
        rb = p.get("rowbias")
        y = ops.linear(ops.cast_bf16(lat), p["w_out"], bias=p["b_out"], rowbias=rb.repeat(B, 1) if rb is not None else None,
                       rows_per_group=(Lq // self.video_length) if rb is not None else 0)
        return ops.layernorm(y, p["g_out"], p["be_out"], want_f32=True)[-1].view(B, Lq, -1)
