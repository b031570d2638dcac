"""Algorithmic FLOP model of one UNet pass (the figure `roofline.achieved` is computed from).

Counting rules of SURVEY.md §8d / BASELINE.md §3: GEMM and convolution = 2*M*N*K; attention = 4*Lq*Lk*d*heads
counted DENSE (the epipolar mask is ignored); the text / image context K,V projections are counted as the
reference executes them — once per frame — even though this implementation computes them once per sample.
Checked against the reference's own FlopCounterMode totals: cond 7.875 TFLOP, uncond 7.121 TFLOP (B=1).
"""
from __future__ import annotations

from typing import Dict, Optional, Set

from .config import Layer, UNetConfig, build_topology


def unet_pass_flops(cfg: UNetConfig, B: int, hw: int, ctx_len: int, per_frame_ctx: bool, camera: bool = True,
                    only_blocks: Optional[Set[str]] = None) -> Dict[str, float]:
    """Returns a breakdown in FLOPs.  hw = latent height = width; ctx_len = tokens of the cross-attention context.
    only_blocks: restrict the count to the named blocks ("input_blocks.1", "init_attn", "middle_block", ...)."""
    T = cfg.temporal_length
    topo = build_topology(cfg)
    ted = cfg.time_embed_dim
    D = cfg.context_dim
    n_text = cfg.text_context_len
    n_img = 16 if per_frame_ctx else ctx_len - n_text
    out = dict(conv=0.0, linear=0.0, ctx_kv=0.0, attn=0.0, epipolar=0.0)

    def lin(M, K, N):
        return 2.0 * M * K * N

    def res(L: Layer, H):
        M = B * T * H * H
        f = lin(M, 9 * L.cin, L.cout) + lin(M, 9 * L.cout, L.cout)
        if L.cin != L.cout:
            f += lin(M, L.cin, L.cout)
        f += 4 * lin(M, 3 * L.cout, L.cout)
        out["conv"] += f
        out["linear"] += lin(B * T, ted, L.cout)

    def ff(M, C):
        out["linear"] += lin(M, C, 8 * C) + lin(M, 4 * C, C)

    def spatial(L: Layer, H):
        C, M, hwn = L.cin, B * T * H * H, H * H
        out["linear"] += 2 * lin(M, C, C)                      # proj_in / proj_out
        out["linear"] += 3 * lin(M, C, C) + lin(M, C, C)       # attn1 q,k,v + to_out
        out["attn"] += 4.0 * B * T * hwn * hwn * C
        out["linear"] += lin(M, C, C) + lin(M, C, C)           # attn2 q + to_out
        out["ctx_kv"] += 2 * lin(B * T * n_text, D, C) + 2 * lin(B * T * n_img, D, C)
        out["attn"] += 4.0 * B * T * hwn * (n_text + n_img) * C
        ff(M, C)

    def temporal(L: Layer, H):
        C, M = L.cin, B * T * H * H
        inner = L.heads * 64
        out["linear"] += lin(M, C, inner) + lin(M, inner, C)   # proj_in / proj_out
        out["linear"] += 2 * (3 * lin(M, inner, inner) + lin(M, inner, inner))
        out["attn"] += 2 * 4.0 * B * H * H * T * T * inner
        ff(M, inner)
        if camera and L.epipolar:
            Lq = T * H * H
            out["linear"] += lin(M, inner, inner)              # pluker_projection
            out["linear"] += 3 * lin(M, inner, inner) + lin(M, inner, inner)
            out["linear"] += 2 * lin(B * cfg.num_register_tokens, inner, inner)
            out["epipolar"] += 4.0 * B * Lq * (Lq + cfg.num_register_tokens) * inner

    def run(L: Layer, H):
        if L.kind == "conv_in":
            out["conv"] += lin(B * T * H * H, 9 * L.cin, L.cout)
        elif L.kind == "res":
            res(L, H)
        elif L.kind == "spatial":
            spatial(L, H)
        elif L.kind == "temporal":
            temporal(L, H)
        elif L.kind == "down":
            out["conv"] += lin(B * T * (H // 2) * (H // 2), 9 * L.cin, L.cout)
            return H // 2
        elif L.kind == "up":
            out["conv"] += lin(B * T * 4 * H * H, 9 * L.cin, L.cout)
            return H * 2
        return H

    def run_block(name, layers, H):
        if only_blocks is not None and name not in only_blocks:
            snap = dict(out)
            for L in layers:
                H = run(L, H)
            out.update(snap)
            return H
        for L in layers:
            H = run(L, H)
        return H

    H = hw
    if only_blocks is None:
        out["linear"] += 2 * (lin(B, cfg.model_channels, ted) + lin(B, ted, ted))
    for i, blk in enumerate(topo.input_blocks):
        H = run_block(blk.name, blk.layers, H)
        if i == 0:
            run_block("init_attn", [topo.init_attn], H)
    H = run_block("middle_block", topo.middle.layers, H)
    for blk in topo.output_blocks:
        H = run_block(blk.name, blk.layers, H)
    if only_blocks is None:
        out["conv"] += lin(B * T * H * H, 9 * cfg.model_channels, cfg.out_channels)
    out["total"] = sum(out.values())
    return out


def cfg_step_flops(cfg: UNetConfig, B: int, hw: int, n_ctx_frames: int = 2, camera: bool = True) -> float:
    """One classifier-free-guidance DDIM step = cond pass (77 + 256*(1+n) tokens, broadcast; with n = 0 the per-frame rule of
    modified_forwards.py:37-44 applies) + uncond pass (77 + 256, per frame).  camera=False drops the epipolar attention
    (CameraCtrl / MotionCtrl baselines; their small cc_projection linears are not counted)."""
    cond = unet_pass_flops(cfg, B, hw, cfg.text_context_len + 256 * (1 + n_ctx_frames), per_frame_ctx=(n_ctx_frames == 0),
                           camera=camera)["total"]
    unc = unet_pass_flops(cfg, B, hw, cfg.text_context_len + 256, per_frame_ctx=True, camera=camera)["total"]
    return cond + unc
