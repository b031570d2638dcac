"""Torch-tensor-facing wrappers of the C ABI (include/camc2v_b200.h).

PyTorch is used here for device memory and streams only: every wrapper allocates its output with
`torch.empty`, passes raw device pointers + the current CUDA stream to libcamc2v_b200.so and returns.
There is no fallback path: a missing library or an unsupported shape raises `C2VError`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import A_CONV2D, A_CONVT, A_PLAIN, EPI_GEGLU, EPI_GELU, EPI_LINEAR, AttnDesc, GemmDesc

BF16 = _lib.operand_torch_dtype()      # the library's 16-bit operand type: bfloat16 (default build) or float16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda:
        raise _lib.C2VError(f"{name}: expected CUDA {dtype}, got {t.device} {t.dtype}")


# ------------------------------------------------------------------------------------------------ GEMM family
# Split-K partials go through an fp32 workspace (L2-resident at these sizes) + a deterministic reduce kernel.  (A variant that
# reduced inside a thread-block cluster through distributed shared memory measured 10-20 % slower in round 1 and was removed.)


def tile_n(N: int, epi: int = EPI_LINEAR) -> int:
    return _lib.load().c2v_gemm_tile_n(N, epi)


def _gemm(a, w, M, N, Cin, taps, a_mode, nb, d1, d2, lda, bias, rowbias, rows_per_group, residual, out, out_dtype, epi):
    _chk(a, BF16, "gemm.a")
    _chk(w, BF16, "gemm.w")
    n_out = N // 2 if epi == EPI_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype)
    d = GemmDesc()
    d.a, d.w, d.bias, d.rowbias, d.residual, d.out = _p(a), _p(w), _p(bias), _p(rowbias), _p(residual), _p(out)
    d.M, d.N, d.Cin, d.taps = M, N, Cin, taps
    d.a_mode, d.nb, d.d1, d.d2, d.lda = a_mode, nb, d1, d2, lda
    d.rows_per_group = rows_per_group
    d.ldr = residual.stride(0) if residual is not None else 0
    d.ldo = out.stride(0)
    d.out_bf16 = 1 if out.dtype == BF16 else 0
    d.epi = epi
    sk = _lib.load().c2v_gemm_splitk(M, N, Cin, taps, epi)
    if sk > 1:
        d.splitk = sk
        ws = torch.empty((sk, M, N), device=a.device, dtype=F32)     # partial tiles; reduced by the second kernel of the call
        d.ws = _p(ws)
        _lib.LAUNCHES += 1
    _lib.call("c2v_gemm", C.byref(d), _stream())
    return out


def linear(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, out=None, out_dtype=F32, rowbias=None, rows_per_group=0,
           gelu: bool = False):
    """a bf16 [M, K] (row stride may exceed K), w bf16 [N, K] -> [M, N] (+bias +rowbias +residual); gelu=True applies the exact
    erf GELU to the result (the FeedForward of the adaptor, resampler.py:31-38)."""
    M, K = a.shape
    if a.stride(1) != 1 or w.stride(1) != 1 or w.stride(0) != K:
        raise _lib.C2VError("linear: operands must be K-contiguous")
    return _gemm(a, w, M, w.shape[0], K, 1, A_PLAIN, 0, 0, 0, a.stride(0), bias, rowbias, rows_per_group, residual, out, out_dtype,
                 EPI_GELU if gelu else EPI_LINEAR)


def geglu_linear(a: torch.Tensor, w_il: torch.Tensor, bias_il: torch.Tensor):
    """GEGLU projection with the gate fused into the epilogue; w_il / bias_il from geglu_interleave()."""
    M, K = a.shape
    return _gemm(a, w_il, M, w_il.shape[0], K, 1, A_PLAIN, 0, 0, 0, a.stride(0), bias_il, None, 0, None, None, BF16, EPI_GEGLU)


def geglu_interleave(w: torch.Tensor, b: torch.Tensor):
    """Reorder GEGLU.proj rows ([value; gate], attention.py:434-438) so that every N tile of the GEMM holds
    the value columns and the matching gate columns of one output block."""
    N = w.shape[0]
    bn = tile_n(N, EPI_GEGLU)
    if bn <= 0:
        raise _lib.C2VError(f"geglu_interleave: unsupported width {N}")
    half = bn // 2
    inner = N // 2
    idx = torch.arange(N, device=w.device)
    tile, r = idx // bn, idx % bn
    src = torch.where(r < half, tile * half + r, inner + tile * half + (r - half))
    return w[src].contiguous(), b[src].contiguous()


def conv3x3(a: torch.Tensor, w: torch.Tensor, NB: int, H: int, W: int, bias=None, rowbias=None, rows_per_group=0, residual=None,
            out_dtype=F32):
    """a bf16 channels-last [NB*H*W, Cin]; w bf16 [Cout, 9*Cin] (tap-major) -> [NB*H*W, Cout].  pad 1, stride 1."""
    Cin = a.shape[1]
    return _gemm(a, w, NB * H * W, w.shape[0], Cin, 9, A_CONV2D, NB, W, H, Cin, bias, rowbias, rows_per_group, residual, None, out_dtype,
                 EPI_LINEAR)


def conv_t3(a: torch.Tensor, w: torch.Tensor, B: int, T: int, HW: int, bias=None, residual=None, out_dtype=F32):
    """Temporal (3,1,1) convolution: a bf16 [B*T*HW, Cin]; w bf16 [Cout, 3*Cin] -> [B*T*HW, Cout]."""
    Cin = a.shape[1]
    return _gemm(a, w, B * T * HW, w.shape[0], Cin, 3, A_CONVT, B, HW, T, Cin, bias, None, 0, residual, None, out_dtype, EPI_LINEAR)


def copy_rows(src: torch.Tensor, dst: torch.Tensor):
    """dst[r, :] = src[r, :] for 16-bit operand tensors: src [rows, C] contiguous, dst a [rows, C] view whose rows may be strided
    (e.g. a row slice of a wider / longer buffer: concatenation along the token axis without a torch.cat temporary)."""
    _chk(src, BF16, "copy_rows.src")
    _chk(dst, BF16, "copy_rows.dst")
    rows, C_ = src.shape
    if not src.is_contiguous() or tuple(dst.shape) != (rows, C_) or dst.stride(1) != 1:
        raise _lib.C2VError("copy_rows: src must be contiguous [rows, C], dst a row-strided view of the same shape")
    _lib.call("c2v_copy_rows", _p(src), _p(dst), rows, C_, 1, 0, dst.stride(0), _stream())
    return dst


def skinny_linear(x: torch.Tensor, w: torch.Tensor, bias, silu_in: bool):
    _chk(x, F32, "skinny_linear.x")
    _chk(w, BF16, "skinny_linear.w")
    M, K = x.shape
    out = torch.empty((M, w.shape[0]), device=x.device, dtype=F32)
    _lib.call("c2v_skinny_linear", _p(x), _p(w), _p(bias), _p(out), M, w.shape[0], K, int(silu_in), _stream())
    return out


def timestep_embedding(t: torch.Tensor, dim: int):
    t = t.to(torch.int64).contiguous()
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=F32)
    _lib.call("c2v_timestep_embedding", _p(t), _p(out), t.shape[0], dim, _stream())
    return out


# ------------------------------------------------------------------------------------------------ norms
def groupnorm(x: torch.Tensor, gamma, beta, ns: int, rows: int, eps: float, silu: bool):
    """x fp32 [ns*rows, C] channels-last -> bf16 [ns*rows, C]; statistics per (sample, group of C/32)."""
    _chk(x, F32, "groupnorm.x")
    C_ = x.shape[1]
    out = torch.empty((ns * rows, C_), device=x.device, dtype=BF16)
    ws = torch.empty((_lib.load().c2v_groupnorm_ws_floats(ns, rows, C_),), device=x.device, dtype=F32)
    _lib.call("c2v_groupnorm_silu", _p(x), _p(gamma), _p(beta), _p(out), _p(ws), ns, rows, C_, float(eps), int(silu), _stream())
    return out


def softmax_rows(x: torch.Tensor, scale: float) -> torch.Tensor:
    """softmax(scale * x) over the last dim: fp32 [rows, n] -> 16-bit operands [rows, n]."""
    _chk(x, F32, "softmax_rows.x")
    if not x.is_contiguous() or x.dim() != 2:
        raise _lib.C2VError("softmax_rows: x must be a contiguous [rows, n] matrix")
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    _lib.call("c2v_softmax_rows", _p(x), _p(out), x.shape[0], x.shape[1], float(scale), _stream())
    return out


def layernorm(x: torch.Tensor, gamma, beta, add: Optional[torch.Tensor] = None, eps: float = 1e-5, want_f32: bool = False,
              out2: Optional[torch.Tensor] = None):
    """Returns LN(x) in bf16; with `add` also LN(x)+add (bf16; written into `out2` when given: a [rows, C] view whose row stride
    may exceed C); with want_f32 also the un-rounded fp32 LN(x) (last)."""
    _chk(x, F32, "layernorm.x")
    rows, C_ = x.shape
    out = torch.empty((rows, C_), device=x.device, dtype=BF16)
    if add is None:
        out2 = None
    elif out2 is None:
        out2 = torch.empty_like(out)
    else:
        _chk(out2, BF16, "layernorm.out2")
        if tuple(out2.shape) != (rows, C_) or out2.stride(1) != 1:
            raise _lib.C2VError("layernorm.out2 must be a [rows, C] view with unit column stride")
    of = torch.empty((rows, C_), device=x.device, dtype=F32) if want_f32 else None
    _lib.call("c2v_layernorm", _p(x), _p(gamma), _p(beta), _p(out), _p(add), _p(out2), _p(of), rows, C_, float(eps),
              out2.stride(0) if out2 is not None else 0, _stream())
    res = (out,) + ((out2,) if add is not None else ()) + ((of,) if want_f32 else ())
    return res if len(res) > 1 else out


# ------------------------------------------------------------------------------------------------ attention
def attention(q, k, v, bq: int, lq: int, lk: int, heads: int, kv_div: int = 1, out=None, out_scale: float = 1.0, accumulate: bool = False,
              k2=None, v2=None, epi_F=None, epi_grid=None, epi_d: int = 0, mask=None, epi_tile_map=None, epi_bitmask=None):
    """q [bq*lq, >=heads*64] bf16 (row-strided view allowed), k/v [bk*lk, ...]; returns bf16 [bq*lq, heads*64]."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, BF16, "attention." + n)
    hd = heads * 64
    if out is None:
        out = torch.empty((bq * lq, hd), device=q.device, dtype=BF16)
    d = AttnDesc()
    d.q, d.k, d.v, d.out = _p(q), _p(k), _p(v), _p(out)
    d.bq, d.lq, d.lk, d.heads = bq, lq, lk, heads
    d.ldq, d.ldk, d.ldv, d.ldo = q.stride(0), k.stride(0), v.stride(0), out.stride(0)
    d.q_bstride, d.k_bstride, d.v_bstride, d.o_bstride = lq * q.stride(0), lk * k.stride(0), lk * v.stride(0), lq * out.stride(0)
    d.kv_div = kv_div
    d.out_scale = float(out_scale)
    d.accumulate = int(accumulate)
    if k2 is not None:
        d.k2, d.v2, d.lk2, d.ldk2, d.ldv2 = _p(k2), _p(v2), k2.shape[0], k2.stride(0), v2.stride(0)
    if epi_F is not None:
        _chk(epi_F, F32, "attention.epi_F")
        T, H, W = epi_grid
        d.epi_F, d.epi_T, d.epi_H, d.epi_W, d.epi_d = _p(epi_F), T, H, W, epi_d
        if epi_tile_map is not None:
            d.epi_tile_map = _p(epi_tile_map)
        if epi_bitmask is not None:
            d.epi_bitmask = _p(epi_bitmask)
    if mask is not None:
        if mask.dtype not in (torch.bool, torch.uint8) or not mask.is_contiguous():
            raise _lib.C2VError("attention.mask must be contiguous bool/uint8 [bq, lq, lk]")
        d.mask, d.mask_bstride = _p(mask), lq * lk
    _lib.call("c2v_attention", C.byref(d), _stream())
    return out


def attention_temporal(qkv: torch.Tensor, B: int, T: int, HW: int, heads: int, out: Optional[torch.Tensor] = None):
    """qkv bf16 [B*T*HW, 3*heads*64] (q | k | v) -> bf16 [B*T*HW, heads*64]; attention over T per (b, pixel).  `out` may be a
    column block of a wider buffer (row stride > heads*64)."""
    _chk(qkv, BF16, "attention_temporal.qkv")
    if out is None:
        out = torch.empty((B * T * HW, heads * 64), device=qkv.device, dtype=BF16)
    else:
        _chk(out, BF16, "attention_temporal.out")
        if tuple(out.shape) != (B * T * HW, heads * 64) or out.stride(1) != 1:
            raise _lib.C2VError("attention_temporal.out must be a [B*T*HW, heads*64] view with unit column stride")
    _lib.call("c2v_attention_temporal", _p(qkv), _p(out), B, T, HW, heads, out.stride(0), _stream())
    return out


def attention_temporal_hd(qkv: torch.Tensor, B: int, T: int, HW: int, heads: int, head_dim: int):
    """qkv [B*T*HW, 3*heads*head_dim] (q | k | v) -> [B*T*HW, heads*head_dim]; any head dim % 8 == 0, <= 160 (pose encoder)."""
    _chk(qkv, BF16, "attention_temporal_hd.qkv")
    out = torch.empty((B * T * HW, heads * head_dim), device=qkv.device, dtype=BF16)
    _lib.call("c2v_attention_temporal_hd", _p(qkv), _p(out), B, T, HW, heads, head_dim, _stream())
    return out


def pixel_unshuffle_cl(x: torch.Tensor, r: int):
    """fp32 [B, C, T, H, W] -> 16-bit rows [(b, t, y, x), C*r*r] (rearrange + nn.PixelUnshuffle(r) + channel-last)."""
    _chk(x, F32, "pixel_unshuffle_cl.x")
    B, C_, T, H, W = x.shape
    out = torch.empty((B * T * (H // r) * (W // r), C_ * r * r), device=x.device, dtype=BF16)
    _lib.call("c2v_pixel_unshuffle_cl", _p(x), _p(out), B, C_, T, H, W, r, _stream())
    return out


def avgpool2_cl(x: torch.Tensor, N: int, H: int, W: int, want_16: bool = True):
    """fp32 rows [N*H*W, C] -> (fp32 [N*H/2*W/2, C], the same in the operand dtype or None): nn.AvgPool2d(2, 2)."""
    _chk(x, F32, "avgpool2_cl.x")
    C_ = x.shape[1]
    out = torch.empty((N * (H // 2) * (W // 2), C_), device=x.device, dtype=F32)
    o16 = torch.empty(out.shape, device=x.device, dtype=BF16) if want_16 else None
    _lib.call("c2v_avgpool2_cl", _p(x), _p(out), _p(o16), N, H, W, C_, _stream())
    return out, o16


def relu_(x: torch.Tensor):
    _chk(x, BF16, "relu_.x")
    _lib.call("c2v_relu", _p(x), x.numel(), _stream())
    return x


# ------------------------------------------------------------------------------------------------ camera
def epipolar_mask(F: torch.Tensor, H: int, W: int, d: int) -> torch.Tensor:
    """F fp32 [B,T1,T2,3,3] -> bool [B, T1*H*W, T2*H*W] (camcontexti2v.py:202-271), bit-exact.  T1 = T2 for the UNet's temporal
    blocks; T1 = 16 target frames x T2 = 1 + n context frames for the adaptor's conditional mask (camcontexti2v.py:493-521)."""
    _chk(F, F32, "epipolar_mask.F")
    F = F.contiguous()
    B, T1, T2 = F.shape[0], F.shape[1], F.shape[2]
    out = torch.empty((B, T1 * H * W, T2 * H * W), device=F.device, dtype=torch.uint8)
    _lib.call("c2v_epipolar_mask_rect", _p(F), _p(out), B, T1, T2, H, W, d, _stream())
    return out.view(torch.bool)


def epipolar_bitmask(F: torch.Tensor, T: int, H: int, W: int, d: int, out=None):
    """The epipolar mask packed to bits for attention(..., epi_F=F, epi_bitmask=...): int32 [B, q_tiles, k_chunks, 128]
    (bit i of word [b, qt, c, r] = mask[b, 128 qt + r, 32 c + i]); any grid whose token count is a multiple of 128 (the five
    power-of-two grids of the shipped resolutions have a specialised builder); None otherwise."""
    _chk(F, F32, "epipolar_bitmask.F")
    L = T * H * W
    if L % 128:
        return None
    B = F.shape[0]
    if out is None:
        out = torch.empty((B, L // 128, L // 32, 128), device=F.device, dtype=torch.int32)
    assert out.numel() == B * _lib.load().c2v_epipolar_bitmask_words(T, H, W)
    _lib.call("c2v_epipolar_bitmask", _p(F.contiguous()), _p(out), B, T, H, W, d, _stream())
    return out


def epipolar_tile_map(F: torch.Tensor, T: int, H: int, W: int, d: int, out=None):
    """Tile-occupancy bitmap for attention(..., epi_F=F, epi_tile_map=...); None when the grid has no fast path."""
    _chk(F, F32, "epipolar_tile_map.F")
    if H != W or (W, d) not in ((32, 8), (16, 16), (8, 32), (16, 8), (8, 16)):
        return None
    B = F.shape[0]
    L = T * H * W
    words = _lib.load().c2v_epipolar_tile_map_words(T, H, W)
    if out is None:
        out = torch.empty((B, (L + 127) // 128, words), device=F.device, dtype=torch.int32)
    _lib.call("c2v_epipolar_tile_map", _p(F.contiguous()), _p(out), B, T, H, W, d, _stream())
    return out


def plucker(K: torch.Tensor, c2w: torch.Tensor, H: int, W: int, mode: str = "plucker") -> torch.Tensor:
    K = K.contiguous().float()
    c2w = c2w.contiguous().float()
    B, T = K.shape[0], K.shape[1]
    out = torch.empty((B, 6, T, H, W), device=K.device, dtype=F32)
    _lib.call("c2v_plucker", _p(K), _p(c2w), _p(out), B, T, H, W, 1 if mode == "plucker" else 0, _stream())
    return out


# ------------------------------------------------------------------------------------------------ layout / glue
def to_channels_last(x: torch.Tensor, B: int, C_: int, S: int, Cpad: Optional[int] = None, dtype=F32, out=None):
    """x fp32 contiguous, viewed as [B, C, S] -> [B*S, Cpad]."""
    _chk(x, F32, "to_channels_last.x")
    Cpad = Cpad or C_
    if out is None:
        out = torch.empty((B * S, Cpad), device=x.device, dtype=dtype)
    _lib.call("c2v_to_channels_last", _p(x), _p(out), B, C_, S, Cpad, 1 if dtype == BF16 else 0, _stream())
    return out


def from_channels_last(x: torch.Tensor, B: int, C_: int, S: int):
    _chk(x, F32, "from_channels_last.x")
    out = torch.empty((B, C_, S), device=x.device, dtype=F32)
    _lib.call("c2v_from_channels_last", _p(x), _p(out), B, C_, S, _stream())
    return out


# 16-bit copies of the UNNORMALISED fp32 residual stream (operand of the ResBlock 1x1 skip convolution) are stored at this power
# of two and the convolution's packed weights carry the inverse: values up to 65504 / RESIDUAL_PRESCALE = 1.6e7 stay finite in the
# IEEE-half build (an unscaled cast overflows at 65504), and the product is unchanged (both scalings are exact).
RESIDUAL_PRESCALE = 2.0 ** -8


def concat_channels(a: torch.Tensor, b: torch.Tensor, want_f32=True, want_bf16=False, scale16: float = 1.0):
    rows, Ca = a.shape
    Cb = b.shape[1]
    of = torch.empty((rows, Ca + Cb), device=a.device, dtype=F32) if want_f32 else None
    ob = torch.empty((rows, Ca + Cb), device=a.device, dtype=BF16) if want_bf16 else None
    _lib.call("c2v_concat_channels_scaled", _p(a), _p(b), _p(of), _p(ob), rows, Ca, Cb, float(scale16), _stream())
    return of, ob


def cast_bf16(x: torch.Tensor, scale: float = 1.0):
    _chk(x, F32, "cast_bf16.x")
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    if scale == 1.0:
        _lib.call("c2v_cast_bf16", _p(x), _p(out), x.numel(), _stream())
    else:
        _lib.call("c2v_cast_bf16_scaled", _p(x), _p(out), x.numel(), float(scale), _stream())
    return out


def upsample2x(x: torch.Tensor, N: int, H: int, W: int):
    C_ = x.shape[1]
    out = torch.empty((N * 4 * H * W, C_), device=x.device, dtype=BF16)
    _lib.call("c2v_upsample2x", _p(x), _p(out), N, H, W, C_, _stream())
    return out


def im2col_s2(x: torch.Tensor, N: int, H: int, W: int, pad_lo: int = 1):
    """Patches of a stride-2 3x3 conv: pad_lo = 1 symmetric padding 1 (UNet Downsample), pad_lo = 0 right / bottom padding only
    (VAE encoder Downsample)."""
    C_ = x.shape[1]
    out = torch.empty((N * (H // 2) * (W // 2), 9 * C_), device=x.device, dtype=BF16)
    _lib.call("c2v_im2col_s2_pad", _p(x), _p(out), N, H, W, C_, pad_lo, _stream())
    return out


def cfg_ddim_update(x, e_cond, e_uncond, noise, scale, guidance_rescale, a_t, a_prev, sigma_t, sqrt_one_minus_at, e_cond_nocam=None,
                    cam_weight: float = 0.0):
    """Fused CFG combine (+ optional camera guidance term, ddim.py:268-280) + guidance rescale + DDIM update (ddim.py:262-346).
    Returns (x_prev, pred_x0)."""
    for t in (x, e_cond, e_uncond, noise) + ((e_cond_nocam,) if e_cond_nocam is not None else ()):
        _chk(t, F32, "cfg_ddim_update")
    B = x.shape[0]
    n = x.numel() // B
    x_prev = torch.empty_like(x)
    pred_x0 = torch.empty_like(x)
    if e_cond_nocam is None:
        _lib.call("c2v_cfg_ddim_update", _p(x.contiguous()), _p(e_cond.contiguous()), _p(e_uncond.contiguous()), _p(noise.contiguous()),
                  _p(x_prev), _p(pred_x0), B, n, float(scale), float(guidance_rescale), float(a_t), float(a_prev), float(sigma_t),
                  float(sqrt_one_minus_at), _stream())
    else:
        _lib.call("c2v_cfg_ddim_update_cam", _p(x.contiguous()), _p(e_cond.contiguous()), _p(e_uncond.contiguous()), _p(e_cond_nocam.contiguous()),
                  _p(noise.contiguous()), _p(x_prev), _p(pred_x0), B, n, float(scale), float(cam_weight), float(guidance_rescale), float(a_t),
                  float(a_prev), float(sigma_t), float(sqrt_one_minus_at), _stream())
    return x_prev, pred_x0
