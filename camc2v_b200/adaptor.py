"""MultiLatentEpipolarAdaptor on the B200 kernels: the once-per-sample module that turns the VAE latents
of the reference frame + n context frames into the per-frame `c_concat` conditioning of the UNet, by 12 layers of
epipolar-masked cross-attention from 16 x h x w learned queries to the (1 + n) x h x w context tokens.

Mirrors R/model/modules/adaptors.py:36-182 (constructor keywords, parameter names and shapes, forward signature) for the
configuration CamContextI2V ships (configs/models/camcontexti2v_256.yaml:141-152: `plucker` input off, `sinusoidal_embedded`
frame embedding, no positional encoding of the context, output_queries = num_queries).  The conditional mask
(compute_conditional_epipolar_mask, R/model/camcontexti2v.py:493-521) comes from `camera.conditional_fundamental_matrices` +
`ops.epipolar_mask` (bit-exact, rectangular) and is consumed by `c2v_attention` in its reference-format mask mode; the two
register tokens of every layer are parameters, so their K / V rows are projected once at weight-pack time.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .modules import _Prepared, _bf16, _f32, _pad_cols, _pad_rows

F32 = torch.float32


class EpipolarCrossAttention(nn.Module):
    """Parameter holder of R/model/modules/epipolar.py:43-72."""

    def __init__(self, query_dim, context_dim=None, out_dim=None, heads=8, dim_head=64, dropout=0.0, num_register_tokens=0):
        super().__init__()
        assert dim_head == 64, "the attention kernel is specialised for head dim 64"
        inner = heads * dim_head
        context_dim = context_dim or query_dim
        self.heads, self.num_register_tokens = heads, num_register_tokens
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, out_dim or query_dim), nn.Dropout(dropout))
        if num_register_tokens > 0:
            self.register_tokens = nn.Parameter(torch.randn((1, num_register_tokens, context_dim)))


def FeedForward(dim, mult=4):
    """R/lvdm/modules/encoders/resampler.py:31-38."""
    inner = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(), nn.Linear(inner, dim, bias=False))


class MultiLatentEpipolarAdaptor(_Prepared):
    def __init__(self, query_dim=512, depth=8, dim_head=64, heads=16, num_queries=1024, output_queries=None, embedding_dim=768, output_dim=1024,
                 ff_mult=4, num_register_tokens=2, use_mask=True, checkpoint=False, video_length=None, use_plucker_embedding=False,
                 allow_plucker_embedding_param=False, context_positional_encoding=False, context_positional_encoding_dim=None,
                 timestep_embedding_type="none", timestep_embedding_dim=32, plucker_embedding_dim=320, plucker_input_strategy="add"):
        super().__init__()
        if use_plucker_embedding or allow_plucker_embedding_param or context_positional_encoding:
            raise NotImplementedError("options the shipped CamContextI2V config leaves off (camcontexti2v_256.yaml:141-152)")
        if output_queries not in (None, num_queries):
            raise NotImplementedError("upscaler (output_queries != num_queries)")
        if timestep_embedding_type not in ("none", "sinusoidal_embedded"):
            raise NotImplementedError(timestep_embedding_type)
        self.num_queries, self.video_length, self.use_mask = num_queries, video_length if video_length is not None else 16, use_mask
        self.timestep_embedding_type, self.timestep_embedding_dim = timestep_embedding_type, timestep_embedding_dim
        self.timestep_embedding_func = None
        if timestep_embedding_type == "sinusoidal_embedded":
            self.timestep_embedding_func = nn.Sequential(nn.Linear(timestep_embedding_dim, query_dim), nn.SiLU(), nn.Linear(query_dim, query_dim))
        n_lat = num_queries * video_length if video_length is not None else num_queries
        self.latents = nn.Parameter(torch.randn(1, n_lat, query_dim) / query_dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, query_dim)
        self.proj_out = nn.Linear(query_dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.plucker_in = None
        # NB: the reference builds the attention with its default 8 heads x 64 (the `heads` / `dim_head` keywords are not forwarded,
        # adaptors.py:101-107) - reproduced, since it fixes the parameter shapes
        self.layers = nn.ModuleList([nn.ModuleList([EpipolarCrossAttention(query_dim=query_dim, context_dim=query_dim, out_dim=query_dim,
                                                                           num_register_tokens=num_register_tokens),
                                                    FeedForward(dim=query_dim, mult=ff_mult)]) for _ in range(depth)])

    # ---------------------------------------------------------------------------------------------- weight packs
    def _prepare(self):
        dev = self.latents.device
        p = {"w_in": _bf16(_pad_cols(self.proj_in.weight.detach(), 64)), "b_in": _f32(self.proj_in.bias),
             "w_out": _bf16(self.proj_out.weight), "b_out": _f32(self.proj_out.bias),
             "g_out": _f32(self.norm_out.weight), "be_out": _f32(self.norm_out.bias), "layers": []}
        for attn, ff in self.layers:
            wkv = _bf16(torch.cat([attn.to_k.weight, attn.to_v.weight], dim=0))
            lp = {"wq": _bf16(attn.to_q.weight), "wkv": wkv, "wo": _bf16(attn.to_out[0].weight), "bo": _f32(attn.to_out[0].bias),
                  "g": _f32(ff[0].weight), "b": _f32(ff[0].bias), "w1": _bf16(ff[1].weight), "w2": _bf16(ff[3].weight), "kv_reg": None}
            if attn.num_register_tokens > 0:      # K / V of the register tokens: constants of the layer (epipolar.py:86-90)
                reg = _bf16(attn.register_tokens[0])
                lp["kv_reg"] = ops.linear(_pad_rows(reg, 8), wkv, out_dtype=ops.BF16)[: attn.num_register_tokens].contiguous()
            p["layers"].append(lp)
        if self.timestep_embedding_func is not None:
            # frame embedding: sinusoid(arange(T)) -> Linear -> SiLU -> Linear, a constant [T, D]; proj_out is linear, so its
            # contribution is folded into a per-frame row bias of the proj_out GEMM:  proj_out(l + e_t) = proj_out(l) + W_out e_t
            t = torch.arange(self.video_length, device=dev, dtype=torch.long)
            e = ops.timestep_embedding(t, self.timestep_embedding_dim)
            f0, f2 = self.timestep_embedding_func[0], self.timestep_embedding_func[2]
            e = ops.skinny_linear(e, _bf16(f0.weight), _f32(f0.bias), False)
            e = ops.skinny_linear(e, _bf16(f2.weight), _f32(f2.bias), True)                 # SiLU on the input of the 2nd linear
            p["rowbias"] = ops.skinny_linear(e, _bf16(self.proj_out.weight), None, False).contiguous()   # [T, output_dim]
        return p

    # ---------------------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, plucker_embedding_features=None) -> torch.Tensor:
        """x [B, C*h*w, embedding_dim] fp32, mask bool/uint8 [B, T*h*w, C*h*w] (True = attend) -> [B, T*h*w, output_dim] fp32."""
        if plucker_embedding_features is not None:
            raise NotImplementedError("plucker input of the adaptor (off in the shipped config)")
        p = self.pk()
        B, Lk, E = x.shape
        Lq, D = self.latents.shape[1], self.latents.shape[2]
        heads = self.layers[0][0].heads
        if not self.use_mask:
            mask = None
        if mask is not None:
            mask = mask.contiguous()
            assert tuple(mask.shape) == (B, Lq, Lk), (mask.shape, (B, Lq, Lk))
        xin = torch.zeros((B * Lk, 64), device=x.device, dtype=ops.BF16)
        xin[:, :E] = _f32(x).view(B * Lk, E)
        ctx = ops.linear(xin, p["w_in"], bias=p["b_in"], out_dtype=ops.BF16)                 # [B*Lk, D]
        lat = _f32(self.latents).expand(B, Lq, D).reshape(B * Lq, D).contiguous()           # fp32 residual stream
        for lp in p["layers"]:
            q = ops.linear(ops.cast_bf16(lat), lp["wq"], out_dtype=ops.BF16)
            kv = ops.linear(ctx, lp["wkv"], out_dtype=ops.BF16)                              # [B*Lk, 2*inner]
            inner = lp["wq"].shape[0]
            kw = {}
            if lp["kv_reg"] is not None:
                kw = dict(k2=lp["kv_reg"][:, :inner], v2=lp["kv_reg"][:, inner:])
            a = ops.attention(q, kv[:, :inner], kv[:, inner:], B, Lq, Lk, heads, mask=mask, **kw)
            lat = ops.linear(a, lp["wo"], bias=lp["bo"], residual=lat)
            n = ops.layernorm(lat, lp["g"], lp["b"])
            h = ops.linear(n, lp["w1"], out_dtype=ops.BF16, gelu=True)
            lat = ops.linear(h, lp["w2"], residual=lat)
        rb = p.get("rowbias")
        y = ops.linear(ops.cast_bf16(lat), p["w_out"], bias=p["b_out"],
                       rowbias=rb.repeat(B, 1) if rb is not None else None, rows_per_group=Lq // self.video_length)
        y = ops.layernorm(y, p["g_out"], p["be_out"], want_f32=True)[-1]
        return y.view(B, Lq, -1)
