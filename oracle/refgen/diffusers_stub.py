"""Stand-in for the two `diffusers` classes the reference's camera_pose_encoder.py imports (:4-5); diffusers is not installed in
this image and not vendored in /root/reference.  TEST INFRASTRUCTURE (golden generation only).  Restated from the published
library: parameter names, shapes and math of `Attention` (default AttnProcessor2_0 path) and `FeedForward(activation_fn="geglu")`.
install() registers them as `diffusers.models.attention_processor.Attention` / `diffusers.models.attention.FeedForward`."""
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class AttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        assert attention_mask is None
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        B, L, _ = hidden_states.shape
        q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
        hd = q.shape[-1] // attn.heads
        q, k, v = (t.view(B, -1, attn.heads, hd).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, attn.heads * hd)
        o = attn.to_out[1](attn.to_out[0](o))
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False, upcast_attention=False,
                 out_bias=True, rescale_output_factor=1.0, **unused):
        super().__init__()
        inner = dim_head * heads
        cross = query_dim if cross_attention_dim is None else cross_attention_dim
        self.heads, self.scale, self.rescale_output_factor = heads, dim_head ** -0.5, rescale_output_factor
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross, inner, bias=bias)
        self.to_v = nn.Linear(cross, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = AttnProcessor2_0()


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", **unused):
        super().__init__()
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


def install():
    if "diffusers" in sys.modules:
        return
    root, models = types.ModuleType("diffusers"), types.ModuleType("diffusers.models")
    ap, at = types.ModuleType("diffusers.models.attention_processor"), types.ModuleType("diffusers.models.attention")
    ap.Attention, at.FeedForward = Attention, FeedForward
    root.models, models.attention_processor, models.attention = models, ap, at
    sys.modules.update({"diffusers": root, "diffusers.models": models, "diffusers.models.attention_processor": ap,
                        "diffusers.models.attention": at})
