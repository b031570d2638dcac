"""Host-side mirror of the reference's module API for the per-step denoising path.

Class names, constructor arguments, `forward` signatures and `state_dict` keys follow the reference
(R = CamContextI2V/):
    UNetModel, ResBlock, TemporalConvBlock, Downsample, Upsample, TimestepEmbedSequential
                                           R/lvdm/modules/networks/openaimodel3d.py:30-624
    CrossAttention, BasicTransformerBlock, SpatialTransformer, TemporalTransformer, GEGLU, FeedForward
                                           R/lvdm/modules/attention.py:44-458
    Epipolar, EpipolarCrossAttention       R/model/modules/epipolar.py:43-157
    camera-conditioned forwards            R/model/modules/modified_forwards.py:29-131, 384-450, 505-536
so a checkpoint of the reference loads unchanged (`load_state_dict(strict=True)`), and the modules can be
swapped in through the reference's own `instantiate_from_config` targets.

The nn.Parameters are only storage.  All arithmetic runs in libcamc2v_b200.so (camc2v_b200.ops): the
modules keep the residual stream as a channels-last fp32 token matrix [B*T*H*W, C] ("CL") and hand bf16
operands to the tcgen05 GEMM / attention kernels.  Spatial ((b t) (h w) c), temporal ((b h w) t c) and
epipolar (b (t h w) c) token orders of the reference are all the SAME buffer here — no rearrange kernels.
The reference-layout `forward(...)` of every module converts at the boundary and calls the CL fast path.
There is no torch fallback for any operator.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .config import Layer, UNetConfig, build_topology

BF16 = _lib.operand_torch_dtype()      # the library's 16-bit operand type: bfloat16 (default build) or float16
F32 = torch.float32


# =================================================================================================
# helpers
# =================================================================================================
@dataclass
class Dims:
    B: int
    T: int
    H: int
    W: int

    @property
    def HW(self):
        return self.H * self.W

    @property
    def BT(self):
        return self.B * self.T

    @property
    def M(self):
        return self.B * self.T * self.H * self.W


@dataclass
class ContextPack:
    """Cross-attention context of one UNet pass, bf16, already split as the reference does per layer
    (attention.py:93-98) and with the frame broadcast of modified_forwards.py:38-44 expressed as kv_div."""
    text: torch.Tensor          # [Bt*77, 1024]
    text_div: int               # query batch (b t) uses text batch (b t) // text_div
    text_len: int
    image: Optional[torch.Tensor]   # [Bi*Li, 1024]
    image_div: int
    image_len: int
    gen: int = 0                    # bumped whenever the token matrices are re-filled in place (invalidates projected K/V)


@dataclass
class CameraLevel:
    pluker: Optional[torch.Tensor]      # CL fp32 [B*T*hw, C]
    F: Optional[torch.Tensor]           # fp32 [B, T, T, 3, 3] (device)
    mask: Optional[torch.Tensor]        # bool [B, L, L] in the reference's format
    d: int                              # downsample factor of this level (origin_h // h)
    add_type: str
    RT: Optional[torch.Tensor] = None   # MotionCtrl: [B, T, 12]


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(BF16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


class GroupNormSpecific(nn.GroupNorm):
    """Parameter holder (R/lvdm/basics.py:78-80).  Statistics are always fp32 in the CUDA kernel."""


class _Prepared(nn.Module):
    """Mixin: device-side bf16 weight packs are built lazily and dropped when parameters are reloaded."""

    def __init__(self):
        super().__init__()
        self._pk: Optional[dict] = None
        self._register_load_state_dict_pre_hook(self._invalidate_hook)

    def _invalidate_hook(self, *a, **k):
        self._pk = None

    def invalidate(self):
        for m in self.modules():
            if isinstance(m, _Prepared):
                m._pk = None

    def _apply(self, fn, recurse=True):
        self._pk = None
        return super()._apply(fn, recurse)

    def pk(self) -> dict:
        if self._pk is None:
            with torch.no_grad():
                self._pk = self._prepare()
        return self._pk

    def _prepare(self) -> dict:  # pragma: no cover - overridden
        return {}


class _RefBindable:
    """The reference's model classes re-bind `forward` / `_forward` of UNet sub-modules BY CLASS NAME when they are constructed
    (camcontexti2v.py:111-170, cami2v.py, cameractrl.py, motionctrl.py: `setattr(module, 'forward', new_forward_for_...)`).
    Those python forwards are what THIS package replaces: the forwards of the classes below already implement the
    camera-conditioned behaviour of modified_forwards.py on the CUDA path.  So that `unet_config.target:
    camc2v_b200.modules.UNetModel` is a drop-in for the UNMODIFIED reference constructor, an attempt to overwrite one of these
    forwards with a plain function is declined (and recorded in `_declined_rebinds` for inspection)."""

    _GUARDED = ("forward", "_forward")

    def __setattr__(self, name, value):
        if name in self._GUARDED and callable(value) and not isinstance(value, nn.Module):
            fn = getattr(value, "__func__", value)
            self.__dict__.setdefault("_declined_rebinds", []).append(getattr(fn, "__name__", repr(fn)))
            return
        super().__setattr__(name, value)


def _conv3x3_pack(conv: nn.Conv2d, pad_cin: int = 0):
    w = conv.weight.detach()
    cout, cin = w.shape[0], w.shape[1]
    if pad_cin and pad_cin > cin:
        w = torch.cat([w, w.new_zeros(cout, pad_cin - cin, 3, 3)], dim=1)
    return _bf16(w.permute(0, 2, 3, 1).reshape(cout, -1)), _f32(conv.bias)


# =================================================================================================
# conv blocks
# =================================================================================================
class Downsample(_Prepared):
    """openaimodel3d.py:51-77 (use_conv=True, dims=2): 3x3 conv, stride 2, pad 1."""

    def __init__(self, channels, use_conv=True, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert use_conv and dims == 2
        self.channels = channels
        self.out_channels = out_channels or channels
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def _prepare(self):
        w, b = _conv3x3_pack(self.op)
        return {"w": w, "b": b}

    def forward_cl(self, h: torch.Tensor, dm: Dims):
        p = self.pk()
        col = ops.im2col_s2(h, dm.BT, dm.H, dm.W)
        return ops.linear(col, p["w"], bias=p["b"]), Dims(dm.B, dm.T, dm.H // 2, dm.W // 2)

    def forward(self, x):
        assert x.shape[1] == self.channels
        n, c, hh, ww = x.shape
        h = ops.to_channels_last(_f32(x), n, c, hh * ww)
        y, dm = self.forward_cl(h, Dims(1, n, hh, ww))
        return ops.from_channels_last(y, n, self.out_channels, dm.HW).view(n, self.out_channels, dm.H, dm.W)


class Upsample(_Prepared):
    """openaimodel3d.py:80-106 (use_conv=True, dims=2): nearest 2x, then 3x3 conv."""

    def __init__(self, channels, use_conv=True, dims=2, out_channels=None, padding=1):
        super().__init__()
        assert use_conv and dims == 2
        self.channels = channels
        self.out_channels = out_channels or channels
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding)

    def _prepare(self):
        w, b = _conv3x3_pack(self.conv)
        return {"w": w, "b": b}

    def forward_cl(self, h: torch.Tensor, dm: Dims):
        p = self.pk()
        up = ops.upsample2x(h, dm.BT, dm.H, dm.W)
        nd = Dims(dm.B, dm.T, dm.H * 2, dm.W * 2)
        return ops.conv3x3(up, p["w"], nd.BT, nd.H, nd.W, bias=p["b"]), nd

    def forward(self, x):
        assert x.shape[1] == self.channels
        n, c, hh, ww = x.shape
        h = ops.to_channels_last(_f32(x), n, c, hh * ww)
        y, dm = self.forward_cl(h, Dims(1, n, hh, ww))
        return ops.from_channels_last(y, n, self.out_channels, dm.HW).view(n, self.out_channels, dm.H, dm.W)


class TemporalConvBlock(_Prepared):
    """openaimodel3d.py:239-279: 4 x [GroupNorm32 + SiLU + Conv3d(3,1,1)] + identity."""

    def __init__(self, in_channels, out_channels=None, dropout=0.0, spatial_aware=False):
        super().__init__()
        assert not spatial_aware
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels

        def conv():
            return nn.Conv3d(out_channels, in_channels, (3, 1, 1), padding=(1, 0, 0))

        self.conv1 = nn.Sequential(nn.GroupNorm(32, in_channels), nn.SiLU(), nn.Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0)))
        self.conv2 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout), conv())
        self.conv3 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout), conv())
        self.conv4 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout), conv())
        nn.init.zeros_(self.conv4[-1].weight)
        nn.init.zeros_(self.conv4[-1].bias)

    def _prepare(self):
        out = {}
        for i, seq in enumerate((self.conv1, self.conv2, self.conv3, self.conv4)):
            gn, conv = seq[0], seq[-1]
            w = conv.weight.detach()
            out[i] = (_f32(gn.weight), _f32(gn.bias), _bf16(w.reshape(w.shape[0], w.shape[1], 3).permute(0, 2, 1).reshape(w.shape[0], -1)),
                      _f32(conv.bias))
        return out

    def forward_cl(self, h: torch.Tensor, dm: Dims):
        p = self.pk()
        y = h
        for i in range(4):
            g, b, w, cb = p[i]
            n = ops.groupnorm(y, g, b, dm.B, dm.T * dm.HW, 1e-5, True)
            y = ops.conv_t3(n, w, dm.B, dm.T, dm.HW, bias=cb, residual=h if i == 3 else None)
        return y

    def forward(self, x):
        b, c, t, hh, ww = x.shape
        h = ops.to_channels_last(_f32(x), b, c, t * hh * ww)
        y = self.forward_cl(h, Dims(b, t, hh, ww))
        return ops.from_channels_last(y, b, c, t * hh * ww).view(b, c, t, hh, ww)


class TimestepBlock(nn.Module):
    pass


BATCH_EMB = os.environ.get("C2V_EMB_BATCH", "1") != "0"      # 0: every ResBlock projects the embedding itself (22 launches per pass)


class EmbPack:
    """Timestep embedding of a pass plus the emb_layers projections of ALL ResBlocks, computed by one skinny GEMM at the top
    of UNetModel.forward (22 launches per pass folded into one): block i reads columns [off, off + Cout) of `all`."""
    __slots__ = ("raw", "all")

    def __init__(self, raw: torch.Tensor, all_: torch.Tensor):
        self.raw, self.all = raw, all_


class ResBlock(_Prepared, TimestepBlock):
    """openaimodel3d.py:109-236 (no scale-shift norm, no up/down, 1x1 skip when channels change)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_scale_shift_norm=False, dims=2, use_checkpoint=False,
                 use_conv=False, up=False, down=False, use_temporal_conv=False, tempspatial_aware=False):
        super().__init__()
        assert not (use_scale_shift_norm or up or down or use_conv) and dims == 2
        self.channels, self.emb_channels = channels, emb_channels
        self.out_channels = out_channels or channels
        self.use_temporal_conv = use_temporal_conv
        self.in_layers = nn.Sequential(GroupNormSpecific(32, channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        out_conv = nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)
        nn.init.zeros_(out_conv.weight)
        nn.init.zeros_(out_conv.bias)
        self.out_layers = nn.Sequential(GroupNormSpecific(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout), out_conv)
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)
        if use_temporal_conv:
            self.temopral_conv = TemporalConvBlock(self.out_channels, self.out_channels, dropout=0.1, spatial_aware=tempspatial_aware)

    def _prepare(self):
        w1, b1 = _conv3x3_pack(self.in_layers[2])
        w2, b2 = _conv3x3_pack(self.out_layers[3])
        p = {"g1": _f32(self.in_layers[0].weight), "be1": _f32(self.in_layers[0].bias), "w1": w1, "b1": b1,
             "g2": _f32(self.out_layers[0].weight), "be2": _f32(self.out_layers[0].bias), "w2": w2, "b2": b2,
             "we": _bf16(self.emb_layers[1].weight), "bemb": _f32(self.emb_layers[1].bias)}
        if not isinstance(self.skip_connection, nn.Identity):
            # the skip convolution reads a 16-bit copy of the UNNORMALISED residual stream stored at RESIDUAL_PRESCALE (fp16 range)
            p["ws"] = _bf16(self.skip_connection.weight.reshape(self.out_channels, self.channels) * (1.0 / ops.RESIDUAL_PRESCALE))
            p["bs"] = _f32(self.skip_connection.bias)
        return p

    def forward_cl(self, h: torch.Tensor, emb: torch.Tensor, dm: Dims, h_bf16: Optional[torch.Tensor] = None):
        """h fp32 CL [M, Cin]; emb fp32 [B, emb_channels] (identical for the T frames of a sample, modified_forwards.py:45)."""
        p = self.pk()
        n = ops.groupnorm(h, p["g1"], p["be1"], dm.BT, dm.HW, 1e-5, True)
        sl = getattr(self, "_emb_slice", None)
        if isinstance(emb, EmbPack) and sl is not None:                             # projected by UNetModel.forward for all blocks at once
            e = emb.all[:, sl[0]:sl[0] + sl[1]]
            if e.shape[0] > 1:
                e = e.contiguous()
        else:
            raw = emb.raw if isinstance(emb, EmbPack) else emb
            e = ops.skinny_linear(raw, p["we"], p["bemb"], True)                    # Linear(SiLU(emb)) -> [B, Cout]
        h1 = ops.conv3x3(n, p["w1"], dm.BT, dm.H, dm.W, bias=p["b1"], rowbias=e, rows_per_group=dm.T * dm.HW)
        n = ops.groupnorm(h1, p["g2"], p["be2"], dm.BT, dm.HW, 1e-5, True)
        if "ws" in p:
            skip = ops.linear(h_bf16 if h_bf16 is not None else ops.cast_bf16(h, ops.RESIDUAL_PRESCALE), p["ws"], bias=p["bs"])
        else:
            skip = h
        h2 = ops.conv3x3(n, p["w2"], dm.BT, dm.H, dm.W, bias=p["b2"], residual=skip)
        if self.use_temporal_conv:
            h2 = self.temopral_conv.forward_cl(h2, dm)
        return h2

    def forward(self, x, emb, batch_size=None):
        n, c, hh, ww = x.shape
        b = batch_size or n
        dm = Dims(b, n // b, hh, ww)
        h = ops.to_channels_last(_f32(x), n, c, hh * ww)
        # reference passes emb already repeated over frames [(b t), E]; one row per sample is enough here
        e = _f32(emb).view(b, n // b, -1)[:, 0].contiguous()
        y = self.forward_cl(h, e, dm)
        return ops.from_channels_last(y, n, self.out_channels, hh * ww).view(n, self.out_channels, hh, ww)


# =================================================================================================
# derived-state caches: eviction that can never pull a buffer from under a captured CUDA graph
# =================================================================================================
_PINNED: Dict[int, int] = {}            # id(cache entry list) -> pin count; a pinned entry is never evicted
_RECORD: Optional[dict] = None          # id -> entry: every cache entry touched since begin_cache_record() (also never evicted)
CACHE_INSERTS = 0                       # number of cache entries created so far (the sampler checks that a capture creates none)


def _touch(ent: list) -> list:
    if _RECORD is not None:
        _RECORD[id(ent)] = ent
    return ent


def _inserted(ent: list) -> list:
    global CACHE_INSERTS
    CACHE_INSERTS += 1
    return _touch(ent)


def _evict(cache: dict, limit: int) -> None:
    """Drop the oldest entries once `cache` holds more than `limit`, except entries pinned by a captured CUDA graph and entries
    touched by the warm-up / capture in progress.  (The round-1 wholesale `.clear()` could free tile maps / packed masks /
    projected K/V that a captured graph still points at; replays then read recycled memory.)"""
    if len(cache) <= limit:
        return
    for k in list(cache.keys()):
        if len(cache) <= limit:
            break
        e = cache[k]
        if id(e) not in _PINNED and not (_RECORD is not None and id(e) in _RECORD):
            del cache[k]


def begin_cache_record() -> None:
    """Start recording the cache entries a sequence of passes touches (the sampler brackets warm-up + capture with this)."""
    global _RECORD
    _RECORD = {}


def end_cache_record() -> list:
    global _RECORD
    ents = list(_RECORD.values()) if _RECORD is not None else []
    _RECORD = None
    return ents


def pin_entries(ents: list) -> list:
    """Pin exactly the cache entries a captured graph uses: they are kept (and still refreshed in place) until
    `unpin_caches(token)`.  The token holds the entries, so their buffers stay alive as long as the graph does."""
    for e in ents:
        _PINNED[id(e)] = _PINNED.get(id(e), 0) + 1
    return list(ents)


def unpin_caches(token: Optional[list]) -> None:
    for e in token or ():
        n = _PINNED.get(id(e), 0) - 1
        if n <= 0:
            _PINNED.pop(id(e), None)
        else:
            _PINNED[id(e)] = n


# =================================================================================================
# attention blocks
# =================================================================================================
class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_Prepared):
    """attention.py:441-458 with glu=True: Linear(C, 8C) -> x * gelu(gate) -> Linear(4C, C)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=True, dropout=0.0):
        super().__init__()
        assert glu
        inner = int(dim * mult)
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim))

    def _prepare(self):
        w, b = ops.geglu_interleave(_bf16(self.net[0].proj.weight), _f32(self.net[0].proj.bias))
        return {"w1": w, "b1": b, "w2": _bf16(self.net[2].weight), "b2": _f32(self.net[2].bias)}

    def forward_cl(self, n_bf16: torch.Tensor, residual: Optional[torch.Tensor], out_dtype=F32):
        p = self.pk()
        u = ops.geglu_linear(n_bf16, p["w1"], p["b1"])
        return ops.linear(u, p["w2"], bias=p["b2"], residual=residual, out_dtype=out_dtype)

    def forward(self, x):
        shp = x.shape
        y = self.forward_cl(ops.cast_bf16(_f32(x).view(-1, shp[-1])), None)
        return y.view(*shp[:-1], -1)


class CrossAttention(_Prepared):
    """attention.py:44-211.  q/k/v Linear without bias, scale 64^-0.5, to_out Linear with bias; optional image
    cross-attention branch (to_k_ip / to_v_ip, learnable gate tanh(alpha)+1)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0, relative_position=False, temporal_length=None,
                 video_length=None, image_cross_attention=False, image_cross_attention_scale=1.0,
                 image_cross_attention_scale_learnable=False, text_context_len=77):
        super().__init__()
        assert dim_head == 64 and not relative_position, "the sm_100a attention kernels are specialised for head dim 64"
        inner = dim_head * heads
        self.context_dim = context_dim
        cdim = context_dim if context_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head = heads, dim_head
        self.temporal_length = temporal_length
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cdim, inner, bias=False)
        self.to_v = nn.Linear(cdim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))
        self.image_cross_attention = image_cross_attention
        self.image_cross_attention_scale = image_cross_attention_scale
        self.image_cross_attention_scale_learnable = image_cross_attention_scale_learnable
        self.text_context_len = text_context_len
        if image_cross_attention:
            self.to_k_ip = nn.Linear(cdim, inner, bias=False)
            self.to_v_ip = nn.Linear(cdim, inner, bias=False)
            if image_cross_attention_scale_learnable:
                self.register_parameter("alpha", nn.Parameter(torch.tensor(0.0)))

    def _prepare(self):
        p = {"wo": _bf16(self.to_out[0].weight), "bo": _f32(self.to_out[0].bias), "wq": _bf16(self.to_q.weight)}
        if self.context_dim is None:
            p["wqkv"] = _bf16(torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0))
        else:
            p["wkv"] = _bf16(torch.cat([self.to_k.weight, self.to_v.weight], dim=0))
            if self.image_cross_attention:
                p["wkv_ip"] = _bf16(torch.cat([self.to_k_ip.weight, self.to_v_ip.weight], dim=0))
                gate = self.image_cross_attention_scale
                if self.image_cross_attention_scale_learnable:
                    gate = gate * (math.tanh(float(self.alpha.detach().float().cpu())) + 1.0)
                p["gate"] = float(gate)
        return p

    # ---- fast paths (inputs already LayerNorm-ed, bf16, CL) ----
    def self_spatial(self, n: torch.Tensor, bq: int, lq: int, residual, out_dtype=F32):
        p = self.pk()
        C = self.heads * 64
        qkv = ops.linear(n, p["wqkv"], out_dtype=BF16)
        o = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], bq, lq, lq, self.heads)
        return ops.linear(o, p["wo"], bias=p["bo"], residual=residual, out_dtype=out_dtype)

    def temporal_heads(self, n: torch.Tensor, dm: Dims, out: Optional[torch.Tensor] = None):
        """The attention output BEFORE to_out (bf16 [B*T*HW, heads*64]); `out` may be a column block of a wider buffer."""
        qkv = ops.linear(n, self.pk()["wqkv"], out_dtype=BF16)
        return ops.attention_temporal(qkv, dm.B, dm.T, dm.HW, self.heads, out=out)

    def self_temporal(self, n: torch.Tensor, dm: Dims, residual, out_dtype=F32):
        p = self.pk()
        return ops.linear(self.temporal_heads(n, dm), p["wo"], bias=p["bo"], residual=residual, out_dtype=out_dtype)

    def _context_kv(self, ctx: "ContextPack", w: torch.Tensor, slot: str):
        """K|V projection of the text / image context tokens.  The context is a per-sample constant (it does not depend on
        the timestep), so the projection is computed once per sample and layer — the reference recomputes it for each of
        the 16 frames, 2 passes and 25 steps.  Cached per token buffer; when the buffer is refilled in place (`ctx.gen`
        changes) the projection is recomputed into the same output buffer, which keeps CUDA-graph pointers valid."""
        src = ctx.text if slot == "text" else ctx.image
        cache = self.__dict__.setdefault("_ctx_cache", {})
        key = (slot, src.data_ptr(), tuple(src.shape))
        ent = cache.get(key)
        if ent is None:
            _evict(cache, 8)
            ent = cache[key] = _inserted([ctx.gen, ops.linear(src, w, out_dtype=BF16), ctx, w, slot])
        elif ent[0] != ctx.gen:
            ent[0] = ctx.gen
            ops.linear(src, w, out=ent[1])
        return _touch(ent)[1]

    def cross(self, n: torch.Tensor, ctx: ContextPack, bq: int, lq: int, residual, out_dtype=F32):
        p = self.pk()
        C = self.heads * 64
        q = ops.linear(n, p["wq"], out_dtype=BF16)
        kv = self._context_kv(ctx, p["wkv"], "text")                   # once per sample, not per frame / pass / step
        o = ops.attention(q, kv[:, :C], kv[:, C:], bq, lq, ctx.text_len, self.heads, kv_div=ctx.text_div)
        if self.image_cross_attention and ctx.image is not None:
            kvi = self._context_kv(ctx, p["wkv_ip"], "image")
            ops.attention(q, kvi[:, :C], kvi[:, C:], bq, lq, ctx.image_len, self.heads, kv_div=ctx.image_div, out=o,
                          out_scale=p["gate"], accumulate=True)
        return ops.linear(o, p["wo"], bias=p["bo"], residual=residual, out_dtype=out_dtype)

    # ---- reference signature ----
    def forward(self, x, context=None, mask=None):
        assert mask is None
        b, nq, c = x.shape
        n = ops.cast_bf16(_f32(x).view(b * nq, c))
        if context is None:
            if self.temporal_length is not None:
                # temporal layout of the reference: x is [(b hw), t, c]
                y = self.self_temporal(_transpose_tokens(n, b, nq), Dims(1, nq, b, 1), None)
                return _transpose_tokens(y, nq, b, F32).view(b, nq, c)
            y = self.self_spatial(n, b, nq, None)
        else:
            y = self.cross(n, make_context_pack(context, 1, self.text_context_len, per_frame=False), b, nq, None)
        return y.view(b, nq, -1)


def _transpose_tokens(x: torch.Tensor, a: int, b: int, dtype=None):
    """[a*b, C] in (a, b) order -> (b, a) order.  Boundary-only helper (standalone module calls)."""
    C = x.shape[1]
    y = x.view(a, b, C).transpose(0, 1).contiguous().view(a * b, C)
    return y if dtype is None else y.to(dtype)


_CONTEXT_CACHE: Dict[tuple, list] = {}


def make_context_pack(context: torch.Tensor, T: int, text_len: int = 77, per_frame: Optional[bool] = None) -> ContextPack:
    """context [B, L, D] as given to UNetModel.forward.  Frame broadcast rule of modified_forwards.py:37-44:
    L == 77 + 16*T -> frame f sees its own 16 image tokens, otherwise every frame sees all image tokens.
    The pack (bf16 text / image token matrices) is cached per context buffer: the context is constant over the sampling
    loop, and stable buffers let every cross-attention layer keep its projected K/V."""
    key = (context.data_ptr(), tuple(context.shape), T, text_len, per_frame, str(context.device))
    ent = _CONTEXT_CACHE.get(key)
    if ent is not None and ent[0] == context._version:
        return _touch(ent)[1]
    pack = _make_context_pack(context, T, text_len, per_frame, ent[1] if ent is not None else None)
    if ent is not None:                                  # same buffer refilled in place: keep the (possibly pinned) entry object
        ent[0], ent[1] = context._version, pack
        return _touch(ent)[1]
    _evict(_CONTEXT_CACHE, 16)
    _CONTEXT_CACHE[key] = _inserted([context._version, pack, context])
    return pack


def refresh_context_caches(model: Optional[nn.Module] = None) -> None:
    """After refilling static context buffers in place (new video, CUDA-graph replay): recast the token matrices and
    re-project every layer's cached context K/V into their existing buffers."""
    for key, ent in list(_CONTEXT_CACHE.items()):
        ctx = ent[2]
        if ent[0] != ctx._version:
            ent[1] = _make_context_pack(ctx, key[2], key[3], key[4], ent[1])
            ent[0] = ctx._version
    if model is not None:
        for m in model.modules():
            cache = m.__dict__.get("_ctx_cache")
            if cache:
                for e in cache.values():
                    if e[0] != e[2].gen:
                        m._context_kv(e[2], e[3], e[4])


def _cast_into(src_f32: torch.Tensor, old: Optional[torch.Tensor]):
    if old is None:
        return ops.cast_bf16(src_f32)
    from . import _lib
    _lib.call("c2v_cast_bf16", src_f32.data_ptr(), old.data_ptr(), src_f32.numel(), torch.cuda.current_stream().cuda_stream)
    return old


def _make_context_pack(context: torch.Tensor, T: int, text_len: int, per_frame: Optional[bool], old: Optional[ContextPack]) -> ContextPack:
    """With `old` (same source buffer refilled in place) the token matrices are re-cast INTO the old pack's buffers and the SAME
    pack object is returned with `gen` bumped: every layer's `_ctx_cache` entry holds that object, so `refresh_context_caches`
    sees the new generation and re-projects K/V into the buffers a captured graph points at."""
    B, L, D = context.shape
    if per_frame is None:
        per_frame = (L == text_len + T * 16)
    ctx = _f32(context)
    text = _cast_into(ctx[:, :text_len].contiguous().view(B * text_len, D), old.text if old is not None else None)
    image, idiv, ilen = None, T, 0
    if L > text_len:
        img = ctx[:, text_len:].contiguous()
        if per_frame:
            ilen, idiv = 16, 1
            image = _cast_into(img.view(B * T * 16, D), old.image if old is not None else None)
        else:
            ilen, idiv = L - text_len, T
            image = _cast_into(img.view(B * ilen, D), old.image if old is not None else None)
    if old is not None:
        old.gen += 1
        return old
    return ContextPack(text, T, text_len, image, idiv, ilen, 0)


class EpipolarCrossAttention(_Prepared):
    """epipolar.py:43-102: q/k/v without bias, learned register tokens prepended to the context before to_k/to_v,
    SDPA with the boolean epipolar mask (registers always visible), to_out with bias."""

    def __init__(self, query_dim, context_dim=None, out_dim=None, heads=8, dim_head=64, dropout=0.0, num_register_tokens=0):
        super().__init__()
        assert dim_head == 64
        inner = dim_head * heads
        cdim = context_dim if context_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cdim, inner, bias=False)
        self.to_v = nn.Linear(cdim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, out_dim if out_dim is not None else query_dim), nn.Dropout(dropout))
        self.num_register_tokens = num_register_tokens
        if num_register_tokens > 0:
            self.register_tokens = nn.Parameter(torch.randn((1, num_register_tokens, cdim)))

    def _prepare(self):
        p = {"wqkv": _bf16(torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0)),
             "wo": _bf16(self.to_out[0].weight), "bo": _f32(self.to_out[0].bias)}
        if self.num_register_tokens > 0:
            # the register tokens are parameters: their keys / values are constants of the layer
            reg = _bf16(self.register_tokens[0])
            wkv = _bf16(torch.cat([self.to_k.weight, self.to_v.weight], dim=0))
            p["kv_reg"] = ops.linear(_pad_rows(reg, 8), wkv, out_dtype=BF16)[: self.num_register_tokens].contiguous()
        return p

    def forward_cl(self, src: torch.Tensor, dm: Dims, cam: CameraLevel, residual, out_dtype=F32):
        """src bf16 CL [B*T*HW, C] (self epipolar attention: context == x, epipolar.py:141-143)."""
        p = self.pk()
        return ops.linear(self.heads_cl(src, dm, cam), p["wo"], bias=p["bo"], residual=residual, out_dtype=out_dtype)

    def heads_cl(self, src: torch.Tensor, dm: Dims, cam: CameraLevel, out: Optional[torch.Tensor] = None):
        """The attention output BEFORE to_out (bf16 [B*T*HW, heads*64]); `src` and `out` may be column blocks of a wider buffer."""
        p = self.pk()
        C = self.heads * 64
        L = dm.T * dm.HW
        qkv = ops.linear(src, p["wqkv"], out_dtype=BF16)
        k2 = v2 = None
        if "kv_reg" in p:
            k2, v2 = p["kv_reg"][:, :C], p["kv_reg"][:, C:]
        kw = {}
        if cam.F is not None:
            kw = dict(epi_F=cam.F, epi_grid=(dm.T, dm.H, dm.W), epi_d=cam.d, epi_tile_map=_tile_map(cam.F, dm.T, dm.H, dm.W, cam.d),
                      epi_bitmask=_bitmask(cam.F, dm.T, dm.H, dm.W, cam.d) if USE_EPI_BITMASK else None)
        elif cam.mask is not None:
            kw = dict(mask=cam.mask)
        return ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], dm.B, L, L, self.heads, k2=k2, v2=v2, out=out, **kw)

    def forward(self, x, context=None, attn_mask=None):
        """Reference signature (efficient_forward): x [B, L1, C], context [B, L2, C], attn_mask bool [B, L1, L2]."""
        B, L1, c = x.shape
        if context is not None and context is not x:
            raise NotImplementedError("cross epipolar attention (MultiLatentEpipolarAdaptor, SURVEY f-1) is outside the per-step path")
        cam = CameraLevel(None, None, attn_mask.contiguous() if attn_mask is not None else None, 0, "")
        y = self.forward_cl(ops.cast_bf16(_f32(x).view(B * L1, c)), Dims(B, L1, 1, 1), cam, None)
        return y.view(B, L1, -1)

    efficient_forward = forward


def _pad_rows(t: torch.Tensor, rows: int):
    if t.shape[0] >= rows:
        return t
    return torch.cat([t, t.new_zeros(rows - t.shape[0], t.shape[1])], dim=0).contiguous()


class Epipolar(nn.Module):
    """epipolar.py:105-157."""

    def __init__(self, query_dim, context_dim, heads, origin_h=256, origin_w=256, is_3d_full_attn=False, num_register_tokens=0,
                 compression_factor=1, attention_resolution=(8, 4, 2, 1), only_on_cond_frame=False, **kwargs):
        super().__init__()
        assert compression_factor == 1 and not only_on_cond_frame
        self.origin_h, self.origin_w = origin_h, origin_w
        self.num_heads = heads
        self.is_3d_full_attn = is_3d_full_attn
        self.epipolar_attn = EpipolarCrossAttention(query_dim=query_dim, context_dim=context_dim, heads=heads,
                                                    dim_head=query_dim // heads, num_register_tokens=num_register_tokens)
        nn.init.zeros_(self.epipolar_attn.to_out[0].weight)
        nn.init.zeros_(self.epipolar_attn.to_out[0].bias)

    @classmethod
    def adopt(cls, ref: nn.Module) -> "Epipolar":
        """Build this class from an instance of the reference's `Epipolar` (R/model/modules/epipolar.py:105-126), which the
        reference's constructors inject with `add_module('epipolar', Epipolar(...))` (camcontexti2v.py:158-166): same
        hyper-parameters, same parameter names / shapes, parameter values copied."""
        attn = ref.epipolar_attn
        new = cls(query_dim=attn.to_q.in_features, context_dim=attn.to_k.in_features, heads=ref.num_heads, origin_h=ref.origin_h,
                  origin_w=ref.origin_w, is_3d_full_attn=ref.is_3d_full_attn, num_register_tokens=attn.num_register_tokens,
                  compression_factor=getattr(ref, "compression_factor", 1), only_on_cond_frame=getattr(ref, "only_on_cond_frame", False))
        new.load_state_dict(ref.state_dict(), strict=True)
        for (_, a), (_, b) in zip(new.named_parameters(), ref.named_parameters()):
            a.requires_grad_(b.requires_grad)
        return new

    def forward(self, features, sample_locs_dict=None, cond_frame_index=None, epipolar_F=None, **kwargs):
        """features [B, T, C, H, W] -> [(B H W), T, C]."""
        B, T, c, H, W = features.shape
        d = self.origin_h // H
        Fm = mask = None
        if not self.is_3d_full_attn:
            if epipolar_F is not None:
                Fm = _f32(epipolar_F).to(features.device)
            elif sample_locs_dict is not None:
                mask = sample_locs_dict.get(d, None)
                mask = mask.to(features.device).contiguous() if mask is not None else None
        x = ops.to_channels_last(_f32(features).view(B * T, c, H * W), B * T, c, H * W)          # [B*T*HW, C], (t, y, x) order
        y = self.epipolar_attn.forward_cl(ops.cast_bf16(x), Dims(B, T, H, W), CameraLevel(None, Fm, mask, d, ""), None)
        return y.view(B, T, H * W, c).permute(0, 2, 1, 3).reshape(B * H * W, T, c)


class BasicTransformerBlock(_RefBindable, _Prepared):
    """attention.py:214-253 plus the camera-conditioned temporal variant (modified_forwards.py:505-536):
        spatial : x = attn1(LN1 x) + x ; x = attn2(LN2 x, context) + x ; x = FF(LN3 x) + x
        temporal: n = LN1 x ; z = pluker_projection(n + p) + Epipolar(n + p) ; x = z + attn1(n) + x ; ...
    """

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True, disable_self_attn=False,
                 attention_cls=None, video_length=None, image_cross_attention=False, image_cross_attention_scale=1.0,
                 image_cross_attention_scale_learnable=False, text_context_len=77, is_output_block=False, ds=1, temporal_length=None):
        super().__init__()
        assert not disable_self_attn and gated_ff
        self.ds = ds
        self.context_dim = context_dim
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout, context_dim=None,
                                    temporal_length=temporal_length)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                                    video_length=video_length, image_cross_attention=image_cross_attention,
                                    image_cross_attention_scale=image_cross_attention_scale,
                                    image_cross_attention_scale_learnable=image_cross_attention_scale_learnable,
                                    text_context_len=text_context_len, temporal_length=temporal_length)
        self.image_cross_attention = image_cross_attention
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.variant = "camcontext"

    def _prepare(self):
        p = {}
        for i, ln in enumerate((self.norm1, self.norm2, self.norm3), 1):
            p[f"g{i}"], p[f"b{i}"] = _f32(ln.weight), _f32(ln.bias)
        for name in ("pluker_projection", "cc_projection"):
            if hasattr(self, name):
                lin = getattr(self, name)
                p[name] = (_bf16(lin.weight), _f32(lin.bias))
        src = self._fused_out_sources()
        if src is not None:
            # z = pluker_projection(n + p) + Epipolar(n + p) ; x = z + attn1(n) + x (modified_forwards.py:519-533) are three
            # Linear layers summed into the stream: one GEMM over the K-concatenated operand [n + p | attn1 heads | epipolar heads]
            # with the weights side by side and the biases added up — one read-modify-write of the fp32 stream instead of three.
            p["w_cat"] = _bf16(torch.cat([l.weight for l in src], dim=1))
            p["b_cat"] = _f32(sum(l.bias.detach().float() for l in src))
            p["cat_ver"] = tuple((id(t), t._version) for l in src for t in (l.weight, l.bias))
        return p

    def _fused_out_sources(self):
        """The three output projections of the camera-conditioned temporal block, when all are present with equal shapes."""
        if not FUSE_TEMPORAL_OUT or self.variant != "camcontext" or not hasattr(self, "epipolar") or not hasattr(self, "pluker_projection"):
            return None
        src = (self.pluker_projection, self.attn1.to_out[0], self.epipolar.epipolar_attn.to_out[0])
        if any(l.weight.shape != src[0].weight.shape or l.bias is None for l in src):
            return None
        return src

    def add_module(self, name, module):
        """The reference injects camera sub-modules into the temporal blocks by name (`epipolar`, `pluker_projection`,
        `cc_projection`; camcontexti2v.py:151-166, cameractrl.py, motionctrl.py).  A reference `Epipolar` instance is adopted as
        this package's class (same parameters); Linear layers are parameter holders as they are.  The block's conditioning
        variant follows from what was injected."""
        if name == "epipolar" and module is not None and not isinstance(module, Epipolar):
            module = Epipolar.adopt(module)
        super().add_module(name, module)
        if name in ("epipolar", "pluker_projection"):
            self.variant = "camcontext"
        elif name == "cc_projection" and module is not None:
            self.variant = "motionctrl" if module.in_features != module.out_features else "cameractrl"
        self._pk = None

    def origin_h(self) -> int:
        """Pixel size of the full frame the epipolar grid refers to (Epipolar.origin_h, epipolar.py:112-113)."""
        return self.epipolar.origin_h if hasattr(self, "epipolar") else 256

    def forward_spatial(self, x: torch.Tensor, ctx: ContextPack, bq: int, lq: int):
        """x fp32 CL [bq*lq, C] -> bf16 (the only consumer is proj_out)."""
        p = self.pk()
        x = self.attn1.self_spatial(ops.layernorm(x, p["g1"], p["b1"]), bq, lq, residual=x)
        x = self.attn2.cross(ops.layernorm(x, p["g2"], p["b2"]), ctx, bq, lq, residual=x)
        return self.ff.forward_cl(ops.layernorm(x, p["g3"], p["b3"]), residual=x, out_dtype=BF16)

    def forward_temporal(self, x: torch.Tensor, dm: Dims, cam: Optional[CameraLevel]):
        p = self.pk()
        has_epi = hasattr(self, "epipolar")
        has_pp = "pluker_projection" in p
        if cam is not None and (has_epi or has_pp) and self.variant == "camcontext":
            if cam.add_type != "add_to_main_branch":
                raise NotImplementedError("only add_type == 'add_to_main_branch' (the shipped configs) is implemented")
            if "w_cat" in p and cam.pluker is not None:
                fs = self._fused_out_sources()
                if p["cat_ver"] != tuple((id(t), t._version) for l in fs for t in (l.weight, l.bias)):
                    self._pk = None                                                 # a child's parameters were rewritten in place
                    p = self.pk()
                C = x.shape[1]
                cat = torch.empty((x.shape[0], 3 * C), device=x.device, dtype=BF16)
                n, src = ops.layernorm(x, p["g1"], p["b1"], add=cam.pluker, out2=cat[:, :C])                # n + p
                self.attn1.temporal_heads(n, dm, out=cat[:, C:2 * C])
                self.epipolar.epipolar_attn.heads_cl(src, dm, cam, out=cat[:, 2 * C:])
                x = ops.linear(cat, p["w_cat"], bias=p["b_cat"], residual=x)        # x + pp(n + p) + attn1(n) + Epipolar(n + p)
                x = self.attn2.self_temporal(ops.layernorm(x, p["g2"], p["b2"]), dm, residual=x)
                return self.ff.forward_cl(ops.layernorm(x, p["g3"], p["b3"]), residual=x, out_dtype=BF16)
            if cam.pluker is not None:
                n, src = ops.layernorm(x, p["g1"], p["b1"], add=cam.pluker)
            else:
                n = src = ops.layernorm(x, p["g1"], p["b1"])
            x = self.attn1.self_temporal(n, dm, residual=x)                         # attn1(n) + x
            if has_pp and cam.pluker is not None:
                w, b = p["pluker_projection"]
                x = ops.linear(src, w, bias=b, residual=x, out=x)                   # + pluker_projection(n + p)
            if has_epi:
                x = self.epipolar.epipolar_attn.forward_cl(src, dm, cam, residual=x)   # + Epipolar(n + p)
        elif cam is not None and self.variant == "cameractrl" and "cc_projection" in p and cam.pluker is not None:
            n, src, nf = ops.layernorm(x, p["g1"], p["b1"], add=cam.pluker, want_f32=True)
            w, b = p["cc_projection"]
            n2 = ops.linear(src, w, bias=b, residual=nf, out_dtype=BF16)            # n + cc_projection(n + p)
            x = self.attn1.self_temporal(n2, dm, residual=x)
        else:
            x = self.attn1.self_temporal(ops.layernorm(x, p["g1"], p["b1"]), dm, residual=x)
        if cam is not None and self.variant == "motionctrl" and "cc_projection" in p and cam.RT is not None:
            x = self._motionctrl_projection(x, dm, cam, p)
        x = self.attn2.self_temporal(ops.layernorm(x, p["g2"], p["b2"]), dm, residual=x)
        return self.ff.forward_cl(ops.layernorm(x, p["g3"], p["b3"]), residual=x, out_dtype=BF16)

    def _motionctrl_projection(self, x, dm, cam, p):
        # x = cc_projection(cat[x, RT12]) (motionctrl_modified_modules.py:184-197): split the (C+12)-wide weight,
        # the 12 pose columns become a per-frame row bias.
        w, b = p["cc_projection"]
        C = x.shape[1]
        if "cc_split" not in p:
            wf = self.cc_projection.weight.detach().float()
            p["cc_split"] = (_bf16(wf[:, :C] * (1.0 / ops.RESIDUAL_PRESCALE)), _bf16(wf[:, C:]))     # x operand stored pre-scaled (fp16 range)
        wx, wrt = p["cc_split"]
        rt = _pad_cols(_f32(cam.RT).view(dm.B * dm.T, -1), 16)
        rb = ops.skinny_linear(rt, _pad_cols(wrt, 16), b, False)                                    # [B*T, C]
        return ops.linear(ops.cast_bf16(x, ops.RESIDUAL_PRESCALE), wx, rowbias=rb, rows_per_group=dm.HW)

    # ---- reference signature ----
    def forward(self, x, context=None, mask=None, camera_condition=None, **kwargs):
        return self._forward(x, context=context, mask=mask, camera_condition=camera_condition)

    def _forward(self, x, context=None, mask=None, camera_condition=None, **kwargs):
        assert mask is None
        b, nq, c = x.shape
        if self.context_dim is not None:
            xs = _f32(x).view(b * nq, c)
            y = self.forward_spatial(xs, make_context_pack(context, 1, per_frame=False), b, nq)
            return y.float().view(b, nq, c)
        # temporal layout of the reference: [(B hw), t, c]
        cam = None
        hw_b = b
        if camera_condition is not None:
            H, W = camera_condition["h"], camera_condition["w"]
            B = b // (H * W)
            cam = camera_level_from_condition(camera_condition, B, nq, H, W, x.device, self.origin_h())
            dm = Dims(B, nq, H, W)
        else:
            dm = Dims(1, nq, hw_b, 1)
        xs = _f32(x).view(dm.B, dm.HW, nq, c).permute(0, 2, 1, 3).reshape(dm.M, c).contiguous()
        y = self.forward_temporal(xs, dm, cam).float()
        return y.view(dm.B, nq, dm.HW, c).permute(0, 2, 1, 3).reshape(b, nq, c)


def _pad_cols(w: torch.Tensor, mult: int):
    k = w.shape[1]
    if k % mult == 0:
        return w.contiguous()
    return torch.cat([w, w.new_zeros(w.shape[0], mult - k % mult)], dim=1).contiguous()


# Derived camera state, cached per source tensor (address + shape).  Both are functions of per-sample constants (F, Pluecker
# features) that do not change over the 25 steps x 2 passes of a sample.  An entry whose source tensor was modified in place
# (torch bumps `_version`) is recomputed INTO THE SAME BUFFER, so pointers captured in a CUDA graph stay valid; callers that
# replay a graph after refilling static conditioning buffers call `refresh_camera_caches()` first.
USE_EPI_BITMASK = os.environ.get("C2V_EPI_BITMASK", "1") != "0"      # packed per-sample mask cache (A/B switch)
FUSE_TEMPORAL_OUT = os.environ.get("C2V_TT_FUSE", "1") != "0"        # one K-concatenated GEMM for the temporal block's three output projections (A/B switch)
_PLUKER_CACHE: Dict[tuple, list] = {}
_TILEMAP_CACHE: Dict[tuple, list] = {}
_BITMASK_CACHE: Dict[tuple, list] = {}


def _tile_map(Fm: torch.Tensor, T: int, H: int, W: int, d: int):
    """Tile-occupancy bitmap of the epipolar mask: a function of F only, so it is built once per sample and level."""
    key = (Fm.data_ptr(), tuple(Fm.shape), T, H, W, d)
    ent = _TILEMAP_CACHE.get(key)
    if ent is None:
        _evict(_TILEMAP_CACHE, 64)
        ent = _TILEMAP_CACHE[key] = _inserted([Fm._version, ops.epipolar_tile_map(Fm, T, H, W, d), Fm])
    elif ent[0] != Fm._version:
        ent[0] = Fm._version
        if ent[1] is not None:
            ops.epipolar_tile_map(Fm, T, H, W, d, out=ent[1])
    return _touch(ent)[1]


def _bitmask(Fm: torch.Tensor, T: int, H: int, W: int, d: int):
    """The epipolar mask packed to 1 bit per pair (32 MB per sample at 32x32x16; the reference's bool mask is 268 MB): built
    once per sample and level, then every layer / pass / step tests a bit instead of re-evaluating the predicate."""
    key = (Fm.data_ptr(), tuple(Fm.shape), T, H, W, d)
    ent = _BITMASK_CACHE.get(key)
    if ent is None:
        _evict(_BITMASK_CACHE, 64)
        ent = _BITMASK_CACHE[key] = _inserted([Fm._version, ops.epipolar_bitmask(Fm, T, H, W, d), Fm])
    elif ent[0] != Fm._version:
        ent[0] = Fm._version
        if ent[1] is not None:
            ops.epipolar_bitmask(Fm, T, H, W, d, out=ent[1])
    return _touch(ent)[1]


def _pluker_cl(p: torch.Tensor) -> torch.Tensor:
    """[B, C, T, h, w] fp32 -> CL fp32 [B*T*hw, C]; constant over the 25 steps x 2 CFG passes, so cached."""
    key = (p.data_ptr(), tuple(p.shape), str(p.device))
    B, C, T, h, w = p.shape
    ent = _PLUKER_CACHE.get(key)
    if ent is None:
        _evict(_PLUKER_CACHE, 64)
        ent = _PLUKER_CACHE[key] = _inserted([p._version, ops.to_channels_last(_f32(p), B, C, T * h * w), p])
    elif ent[0] != p._version:
        ent[0] = p._version
        ops.to_channels_last(_f32(p), B, C, T * h * w, out=ent[1])
    return _touch(ent)[1]


def refresh_camera_caches() -> None:
    """Re-derive (in place) every cached tile map / channels-last Pluecker copy whose source tensor was refilled."""
    for key, ent in list(_TILEMAP_CACHE.items()):
        Fm = ent[2]
        if ent[0] != Fm._version:
            _tile_map(Fm, *key[2:])
    for key, ent in list(_BITMASK_CACHE.items()):
        Fm = ent[2]
        if ent[0] != Fm._version:
            _bitmask(Fm, *key[2:])
    for ent in list(_PLUKER_CACHE.values()):
        if ent[0] != ent[2]._version:
            _pluker_cl(ent[2])


def camera_level_from_condition(cc: dict, B: int, T: int, H: int, W: int, device, origin_h: int = 256) -> CameraLevel:
    """Per-block camera condition in the reference's dict format (modified_forwards.py:66-78) -> kernel inputs.
    `epipolar_F` ([B,T,T,3,3], our addition) selects the in-kernel mask; otherwise `sample_locs_dict[d]` is honoured."""
    pf = cc.get("pluker_embedding_features")
    pl = _pluker_cl(pf.to(device)) if pf is not None else None
    d = origin_h // H
    Fm = cc.get("epipolar_F")
    mask = None
    if Fm is not None:
        Fm = _f32(Fm).to(device)
    else:
        sl = cc.get("sample_locs_dict")
        if sl is not None and sl.get(d) is not None:
            mask = sl[d].to(device).contiguous()
    rt = cc.get("RT")
    return CameraLevel(pl, Fm, mask, d, cc.get("add_type", ""), rt.to(device) if rt is not None else None)


class SpatialTransformer(_Prepared):
    """attention.py:256-320 with use_linear=True."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, use_checkpoint=True, disable_self_attn=False,
                 use_linear=True, video_length=None, image_cross_attention=False, image_cross_attention_scale_learnable=False,
                 is_output_block=False, ds=1):
        super().__init__()
        assert use_linear and depth == 1
        self.ds = ds
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim, video_length=video_length,
                                  image_cross_attention=image_cross_attention,
                                  image_cross_attention_scale_learnable=image_cross_attention_scale_learnable, ds=ds)])
        self.proj_out = nn.Linear(inner, in_channels)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)
        self.use_linear = use_linear

    def _prepare(self):
        return {"g": _f32(self.norm.weight), "b": _f32(self.norm.bias), "wi": _bf16(self.proj_in.weight), "bi": _f32(self.proj_in.bias),
                "wo": _bf16(self.proj_out.weight), "bo": _f32(self.proj_out.bias)}

    def forward_cl(self, h: torch.Tensor, ctx: ContextPack, dm: Dims):
        p = self.pk()
        n = ops.groupnorm(h, p["g"], p["b"], dm.BT, dm.HW, 1e-6, False)
        x = ops.linear(n, p["wi"], bias=p["bi"])
        xb = self.transformer_blocks[0].forward_spatial(x, ctx, dm.BT, dm.HW)
        return ops.linear(xb, p["wo"], bias=p["bo"], residual=h)

    def forward(self, x, context=None, **kwargs):
        n, c, hh, ww = x.shape
        h = ops.to_channels_last(_f32(x), n, c, hh * ww)
        y = self.forward_cl(h, make_context_pack(context, 1, per_frame=False), Dims(1, n, hh, ww))
        return ops.from_channels_last(y, n, c, hh * ww).view(n, c, hh, ww)


class TemporalTransformer(_RefBindable, _Prepared):
    """attention.py:323-428 (only_self_att=True) with the camera-conditioned forward of modified_forwards.py:401-450."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, use_checkpoint=True, use_linear=False,
                 only_self_att=True, causal_attention=False, causal_block_size=1, relative_position=False, temporal_length=None,
                 is_output_block=False, ds=1):
        super().__init__()
        assert only_self_att and not causal_attention and not relative_position and depth == 1
        self.ds = ds
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.use_linear = use_linear
        if use_linear:
            self.proj_in = nn.Linear(in_channels, inner)
            self.proj_out = nn.Linear(inner, in_channels)
        else:   # init_attn (openaimodel3d.py:389-402): Conv1d(k=1), same arithmetic, weight has a trailing unit dim
            self.proj_in = nn.Conv1d(in_channels, inner, kernel_size=1)
            self.proj_out = nn.Conv1d(inner, in_channels, kernel_size=1)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=None, ds=ds, temporal_length=temporal_length)])

    def _prepare(self):
        wi, wo = self.proj_in.weight, self.proj_out.weight
        return {"g": _f32(self.norm.weight), "b": _f32(self.norm.bias), "wi": _bf16(wi.reshape(wi.shape[0], -1)), "bi": _f32(self.proj_in.bias),
                "wo": _bf16(wo.reshape(wo.shape[0], -1)), "bo": _f32(self.proj_out.bias)}

    def forward_cl(self, h: torch.Tensor, dm: Dims, cam: Optional[CameraLevel]):
        p = self.pk()
        n = ops.groupnorm(h, p["g"], p["b"], dm.B, dm.T * dm.HW, 1e-6, False)       # statistics over (t, h, w) per sample
        x = ops.linear(n, p["wi"], bias=p["bi"])
        xb = self.transformer_blocks[0].forward_temporal(x, dm, cam)
        return ops.linear(xb, p["wo"], bias=p["bo"], residual=h)

    def forward(self, x, context=None, camera_condition=None):
        b, c, t, hh, ww = x.shape
        h = ops.to_channels_last(_f32(x), b, c, t * hh * ww)
        cam = None
        if camera_condition is not None:
            cam = camera_level_from_condition(camera_condition, b, t, hh, ww, x.device, self.transformer_blocks[0].origin_h())
        y = self.forward_cl(h, Dims(b, t, hh, ww), cam)
        return ops.from_channels_last(y, b, c, t * hh * ww).view(b, c, t, hh, ww)


class TimestepEmbedSequential(_RefBindable, nn.Sequential, TimestepBlock):
    """openaimodel3d.py:30-48 with the camera_condition argument of modified_forwards.py:384-398."""

    def forward_cl(self, h, emb, ctx, dm: Dims, cam: Optional[CameraLevel], h_bf16=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                h = layer.forward_cl(h, emb, dm, h_bf16)
                h_bf16 = None
            elif isinstance(layer, SpatialTransformer):
                h = layer.forward_cl(h, ctx, dm)
            elif isinstance(layer, TemporalTransformer):
                h = layer.forward_cl(h, dm, cam)
            elif isinstance(layer, (Downsample, Upsample)):
                h, dm = layer.forward_cl(h, dm)
            else:
                raise TypeError(type(layer))
        return h, dm

    def forward(self, x, emb, context=None, batch_size=None, camera_condition=None):
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb, batch_size=batch_size)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            elif isinstance(layer, TemporalTransformer):
                n, c, hh, ww = x.shape
                x5 = x.view(batch_size, n // batch_size, c, hh, ww).permute(0, 2, 1, 3, 4)
                x5 = layer(x5, context, camera_condition=camera_condition)
                x = x5.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)
            else:
                x = layer(x)
        return x


# =================================================================================================
# UNet
# =================================================================================================
class _ConvIn(_Prepared):
    """input_blocks.0.0 (openaimodel3d.py:386): 3x3 conv on 8 channels, zero-padded to one 64-channel K chunk."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1)


class UNetModel(_RefBindable, _Prepared):
    """lvdm 3D-UNet (openaimodel3d.py:281-624) with the camera-conditioned forward (modified_forwards.py:29-131).

    Constructor keywords follow the reference; only the configuration space used by
    configs/models/camcontexti2v_256.yaml and configs/baseline/*.yaml is supported (anything else raises).
    Call `attach_camera_modules()` to add `epipolar` / `pluker_projection` (CamContextI2V, CamI2V) or
    `cc_projection` (CameraCtrl, MotionCtrl) exactly as the reference's model classes do at construction.
    """

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0.0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, context_dim=None, use_scale_shift_norm=False,
                 resblock_updown=False, num_heads=-1, num_head_channels=-1, transformer_depth=1, use_linear=False,
                 use_checkpoint=False, temporal_conv=False, tempspatial_aware=False, temporal_attention=True,
                 use_relative_position=True, use_causal_attention=False, temporal_length=None, use_fp16=False,
                 addition_attention=False, temporal_selfatt_only=True, image_cross_attention=False,
                 image_cross_attention_scale_learnable=False, default_fs=4, fs_condition=False):
        super().__init__()
        if not (dims == 2 and conv_resample and use_linear and temporal_conv and temporal_attention and addition_attention
                and fs_condition and num_head_channels == 64 and transformer_depth == 1 and not use_relative_position
                and not use_causal_attention and not use_scale_shift_norm and not resblock_updown and not tempspatial_aware
                and temporal_selfatt_only):
            raise NotImplementedError("UNetModel: configuration outside the CamContextI2V / DynamiCrafter-256 family")
        self.cfg = UNetConfig(in_channels=in_channels, out_channels=out_channels, model_channels=model_channels,
                              attention_resolutions=tuple(attention_resolutions), num_res_blocks=num_res_blocks,
                              channel_mult=tuple(channel_mult), num_head_channels=num_head_channels, context_dim=context_dim,
                              temporal_length=temporal_length or 16, default_fs=default_fs, epipolar=False, pluker_projection=False)
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.attention_resolutions = list(attention_resolutions)
        self.temporal_length = temporal_length
        self.default_fs = default_fs
        self.fs_condition = fs_condition
        self.addition_attention = addition_attention
        self.dtype = torch.float32
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.fps_embedding = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        nn.init.zeros_(self.fps_embedding[-1].weight)
        nn.init.zeros_(self.fps_embedding[-1].bias)

        topo = build_topology(self.cfg)
        kw_img = dict(image_cross_attention=image_cross_attention, image_cross_attention_scale_learnable=image_cross_attention_scale_learnable)

        def make(L: Layer):
            if L.kind == "conv_in":
                return nn.Conv2d(L.cin, L.cout, 3, padding=1)
            if L.kind == "res":
                return ResBlock(L.cin, ted, dropout, out_channels=L.cout, use_temporal_conv=temporal_conv)
            if L.kind == "spatial":
                return SpatialTransformer(L.cin, L.heads, 64, context_dim=context_dim, use_linear=True, video_length=temporal_length,
                                          ds=L.ds, **kw_img)
            if L.kind == "temporal":
                return TemporalTransformer(L.cin, L.heads, 64, context_dim=context_dim, use_linear=not L.conv_proj,
                                           temporal_length=temporal_length, ds=L.ds)
            if L.kind == "down":
                return Downsample(L.cin, True, out_channels=L.cout)
            if L.kind == "up":
                return Upsample(L.cin, True, out_channels=L.cout)
            raise ValueError(L.kind)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(*[make(L) for L in blk.layers]) for blk in topo.input_blocks])
        self.init_attn = TimestepEmbedSequential(make(topo.init_attn))
        self.middle_block = TimestepEmbedSequential(*[make(L) for L in topo.middle.layers])
        self.output_blocks = nn.ModuleList([TimestepEmbedSequential(*[make(L) for L in blk.layers]) for blk in topo.output_blocks])
        out_conv = nn.Conv2d(model_channels, out_channels, 3, padding=1)
        nn.init.zeros_(out_conv.weight)
        nn.init.zeros_(out_conv.bias)
        self.out = nn.Sequential(GroupNormSpecific(32, topo.out_channels_last), nn.SiLU(), out_conv)
        self.input_ds = [blk.ds for blk in topo.input_blocks]
        self.output_ds = [blk.ds for blk in topo.output_blocks]
        self.middle_ds = topo.middle.ds
        self.origin_h = 256

    # ------------------------------------------------------------------ camera sub-modules
    def temporal_blocks(self, include_init_attn=False):
        init_inner = self.init_attn[0].proj_in.out_channels
        for name, m in self.named_modules():
            if isinstance(m, BasicTransformerBlock) and m.context_dim is None:
                if include_init_attn or m.attn1.to_k.in_features != init_inner:
                    yield name, m

    def attach_camera_modules(self, variant: str = "camcontext", epipolar_config: Optional[dict] = None, pluker_projection: bool = True,
                              pose_dim: int = 12):
        """camcontexti2v.py:111-170 / cami2v.py / cameractrl.py:19-49 / motionctrl.py:19-49."""
        if variant in ("camcontext", "cami2v"):
            ec = dict(origin_h=256, origin_w=256, is_3d_full_attn=False, num_register_tokens=4, attention_resolution=[8, 4, 2, 1],
                      compression_factor=1)
            ec.update(epipolar_config or {})
            self.origin_h = ec["origin_h"]
            for _, m in self.temporal_blocks():
                c = m.attn1.to_k.in_features
                m.variant = "camcontext"
                if pluker_projection:
                    pp = nn.Linear(c, c)
                    nn.init.zeros_(pp.weight)
                    nn.init.zeros_(pp.bias)
                    m.add_module("pluker_projection", pp)
                m.add_module("epipolar", Epipolar(query_dim=c, context_dim=c, heads=m.attn1.heads, **ec))
        elif variant == "cameractrl":
            for _, m in self.temporal_blocks():
                c = m.attn1.to_k.in_features
                m.variant = "cameractrl"
                cc = nn.Linear(c, c)
                nn.init.zeros_(cc.weight)
                nn.init.zeros_(cc.bias)
                m.add_module("cc_projection", cc)
        elif variant == "motionctrl":
            for _, m in self.temporal_blocks(include_init_attn=True):
                c = m.attn2.to_k.in_features
                m.variant = "motionctrl"
                cc = nn.Linear(c + pose_dim, c)
                nn.init.zeros_(cc.weight)
                nn.init.eye_(cc.weight[:c, :c])
                nn.init.zeros_(cc.bias)
                m.add_module("cc_projection", cc)
        else:
            raise ValueError(variant)
        self.invalidate()
        return self

    # ------------------------------------------------------------------ forward
    def _prepare(self):
        w = self.input_blocks[0][0].weight.detach()
        cout, cin = w.shape[0], w.shape[1]
        wpad = torch.cat([w, w.new_zeros(cout, 64 - cin, 3, 3)], dim=1)
        wo, bo = _conv3x3_pack(self.out[2])
        p = {"w_in": _bf16(wpad.permute(0, 2, 3, 1).reshape(cout, -1)), "b_in": _f32(self.input_blocks[0][0].bias),
             "g_out": _f32(self.out[0].weight), "be_out": _f32(self.out[0].bias), "w_out": wo, "b_out": bo}
        for name, seq in (("t", self.time_embed), ("f", self.fps_embedding)):
            p[name] = (_bf16(seq[0].weight), _f32(seq[0].bias), _bf16(seq[2].weight), _f32(seq[2].bias))
        # emb_layers of every ResBlock (openaimodel3d.py:161-167) stacked into one [sum Cout, 4*mc] matrix
        blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        off = 0
        for blk in blocks:
            blk._emb_slice = (off, blk.out_channels)
            off += blk.out_channels
        p["we_all"] = _bf16(torch.cat([blk.emb_layers[1].weight.detach() for blk in blocks], dim=0))
        p["bemb_all"] = _f32(torch.cat([blk.emb_layers[1].bias.detach() for blk in blocks], dim=0))
        return p

    def _embed(self, seq_key: str, idx: torch.Tensor):
        w0, b0, w2, b2 = self.pk()[seq_key]
        e = ops.timestep_embedding(idx, self.model_channels)
        return ops.skinny_linear(ops.skinny_linear(e, w0, b0, False), w2, b2, True)

    def _camera_level(self, cc: Optional[dict], ds: int, dm: Dims, device, middle=False) -> Optional[CameraLevel]:
        """Feature-level routing of modified_forwards.py:66-78 (inputs/outputs) and :92-102 (middle)."""
        if cc is None:
            return None
        local = dict(cc)
        pf = cc.get("pluker_embedding_features")
        if pf is not None:
            if middle:
                local["pluker_embedding_features"] = pf[-1]
            elif ds in self.attention_resolutions:
                local["pluker_embedding_features"] = pf[int(math.log2(ds))]
            else:
                local["pluker_embedding_features"] = None
        return camera_level_from_condition(local, dm.B, dm.T, dm.H, dm.W, device, self.origin_h)

    @torch.no_grad()
    def forward(self, x, timesteps, context=None, features_adapter=None, fs=None, camera_condition=None, **kwargs):
        """x [B, in_channels, T, H, W], timesteps [B], context [B, L, context_dim], fs [B] long -> [B, out_channels, T, H, W]."""
        assert features_adapter is None
        p = self.pk()
        dev = x.device
        b, cin, t, hh, ww = x.shape
        emb = self._embed("t", timesteps.to(dev))
        if self.fs_condition:
            if fs is None:
                fs = torch.full((b,), self.default_fs, dtype=torch.long, device=dev)
            emb = emb + self._embed("f", fs.to(dev))                                  # [B, 4*mc] (tiny; identical for all T frames)
        if BATCH_EMB:
            emb = EmbPack(emb, ops.skinny_linear(emb, p["we_all"], p["bemb_all"], True))
        ctx = make_context_pack(context.to(dev), t)
        dm = Dims(b, t, hh, ww)

        a = ops.to_channels_last(_f32(x), b, cin, t * hh * ww, Cpad=64, dtype=BF16)
        h = ops.conv3x3(a, p["w_in"], dm.BT, dm.H, dm.W, bias=p["b_in"])
        hs: List[tuple] = []
        for i, module in enumerate(self.input_blocks):
            if i > 0:
                cam = self._camera_level(camera_condition, self.input_ds[i], dm, dev)
                h, dm = module.forward_cl(h, emb, ctx, dm, cam)
            if i == 0 and self.addition_attention:
                # no camera condition for init_attn (modified_forwards.py:80-81), except MotionCtrl (motionctrl_modified_modules.py:69)
                cam0 = None
                if camera_condition is not None and self.init_attn[0].transformer_blocks[0].variant == "motionctrl":
                    cam0 = self._camera_level(camera_condition, 1, dm, dev)
                h, dm = self.init_attn.forward_cl(h, emb, ctx, dm, cam0)
            hs.append(h)
        cam = self._camera_level(camera_condition, self.middle_ds, dm, dev, middle=True)
        h, dm = self.middle_block.forward_cl(h, emb, ctx, dm, cam)
        for i, module in enumerate(self.output_blocks):
            hcat, hcat16 = ops.concat_channels(h, hs.pop(), True, True, scale16=ops.RESIDUAL_PRESCALE)   # hcat16 feeds the skip conv only
            cam = self._camera_level(camera_condition, self.output_ds[i], dm, dev)
            h, dm = module.forward_cl(hcat, emb, ctx, dm, cam, hcat16)
        n = ops.groupnorm(h, p["g_out"], p["be_out"], dm.BT, dm.HW, 1e-5, True)
        y = ops.conv3x3(n, p["w_out"], dm.BT, dm.H, dm.W, bias=p["b_out"])            # [M, out_channels]
        return ops.from_channels_last(y, b, self.out_channels, t * hh * ww).view(b, self.out_channels, t, hh, ww)


def build_unet(cfg: UNetConfig = UNetConfig(), variant: Optional[str] = "camcontext") -> UNetModel:
    """UNet of configs/models/camcontexti2v_256.yaml:40-69 (+ camera modules of the chosen variant)."""
    m = UNetModel(in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
                  num_res_blocks=cfg.num_res_blocks, attention_resolutions=list(cfg.attention_resolutions), dropout=0.1,
                  channel_mult=list(cfg.channel_mult), num_head_channels=cfg.num_head_channels, transformer_depth=1,
                  context_dim=cfg.context_dim, use_linear=True, use_checkpoint=False, temporal_conv=True, temporal_attention=True,
                  temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False,
                  temporal_length=cfg.temporal_length, addition_attention=True, image_cross_attention=True,
                  image_cross_attention_scale_learnable=True, default_fs=cfg.default_fs, fs_condition=True)
    if variant and variant != "none":
        m.attach_camera_modules(variant, epipolar_config=dict(origin_h=cfg.origin_h, origin_w=cfg.origin_w,
                                                              num_register_tokens=cfg.num_register_tokens),
                                pluker_projection=cfg.pluker_projection)
    return m.eval()
