"""Kernel-level parity tests (GPU): every C-ABI entry point against a plain fp32 torch evaluation of the
same operator on the same (bf16-rounded) inputs, and the integer/boolean kernels bit-exactly against oracle/.

Tolerances: bf16 operands + fp32 accumulation.  For an output of bf16 dtype the bound is 2^-8 relative to the
row scale (one bf16 rounding) plus accumulation noise; for fp32 outputs 2e-3 relative to max|ref| covers
the fp32 accumulation-order difference of K <= 23k terms of bf16 products.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from camc2v_b200 import ops
    return ops


def rnd(*shape, seed=0, std=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(DEV).to(dtype)


def _dt():
    """16-bit operand dtype of the library under test (bfloat16, or float16 with CAMC2V_B200_OPERANDS=fp16)."""
    from camc2v_b200 import ops
    return ops.BF16


def close(out, ref, tol, what=""):
    out, ref = out.float(), ref.float()
    assert torch.isfinite(out).all(), f"{what}: non-finite output"
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    assert err <= tol * scale, f"{what}: max|err| {err:.4e} > {tol} * max|ref| {scale:.4e}"


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (16384, 320, 320), (1000, 1280, 640), (77, 640, 1024), (4096, 512, 320),
                                   (128, 4, 320), (300, 2560, 1280), (256, 1280, 5120)])
def test_linear(M, N, K):
    ops = _ops()
    a = rnd(M, K, seed=1, dtype=_dt())
    w = rnd(N, K, seed=2, std=K ** -0.5, dtype=_dt())
    bias = rnd(N, seed=3)
    res = rnd(M, N, seed=4)
    ref = a.float() @ w.float().t() + bias + res
    out = ops.linear(a, w, bias=bias, residual=res)
    close(out, ref, 2e-3, "linear fp32")
    out16 = ops.linear(a, w, bias=bias, residual=res, out_dtype=_dt())
    close(out16, ref, 1e-2, "linear bf16")
    # in-place residual (out aliases residual), as used for the attention / FF residual adds
    res2 = res.clone()
    ops.linear(a, w, bias=bias, residual=res2, out=res2)
    close(res2, ref, 2e-3, "linear in-place")


@pytest.mark.parametrize("M,N,K,bf16", [(256, 1280, 5120, False), (1024, 1280, 5120, True), (2048, 640, 2560, False), (200, 1280, 3840, False)])
def test_splitk_is_deterministic_and_needs_a_workspace(M, N, K, bf16):
    """Split-K (partials through the fp32 workspace, fixed summation order in the reduce kernel): correct, run-to-run identical;
    the C entry point rejects a split-K call without a workspace."""
    import ctypes as C
    ops = _ops()
    assert ops._lib.load().c2v_gemm_splitk(M, N, K, 1, 0) > 1
    a = rnd(M, K, seed=1, dtype=_dt())
    w = rnd(N, K, seed=2, std=K ** -0.5, dtype=_dt())
    bias, res, rb = rnd(N, seed=3), rnd(M, N, seed=4), rnd((M + 127) // 128, N, seed=5)
    odt = _dt() if bf16 else torch.float32
    ref = a.float() @ w.float().t() + bias + res + rb.repeat_interleave(128, dim=0)[:M]
    outs = [ops.linear(a, w, bias=bias, residual=res, rowbias=rb, rows_per_group=128, out_dtype=odt) for _ in range(2)]
    close(outs[0], ref, 1e-2 if bf16 else 2e-3, "split-K")
    assert torch.equal(outs[0], outs[1]), "split-K must be deterministic"
    d = ops.GemmDesc()
    out = torch.empty(M, N, device=DEV)
    d.a, d.w, d.out = a.data_ptr(), w.data_ptr(), out.data_ptr()
    d.M, d.N, d.Cin, d.taps, d.lda, d.ldo, d.splitk = M, N, K, 1, K, N, 2
    assert ops._lib.load().c2v_gemm(C.byref(d), None) == 1          # ERR_BAD_ARG: no workspace


def test_linear_strided_a_and_rowbias():
    ops = _ops()
    M, K, N = 512, 320, 640
    big = rnd(M, 3 * K, seed=5, dtype=_dt())
    a = big[:, K:2 * K]
    w = rnd(N, K, seed=6, std=K ** -0.5, dtype=_dt())
    rb = rnd(4, N, seed=7)
    ref = a.float() @ w.float().t() + rb.repeat_interleave(128, dim=0)
    out = ops.linear(a, w, rowbias=rb, rows_per_group=128)
    close(out, ref, 2e-3, "linear strided+rowbias")


@pytest.mark.parametrize("M,N,K,kind", [(16384, 2560, 320, "geglu"), (4096, 5120, 640, "geglu"), (1024, 10240, 1280, "geglu"), (16384, 4096, 512, "geglu"),
                                        (16384, 960, 320, "linear"), (4096, 1920, 640, "linear"), (16384, 1536, 512, "linear"),
                                        (16250, 960, 320, "linear"), (8000, 2560, 320, "geglu"), (16384, 640, 320, "gelu")])
def test_persistent_gemm(M, N, K, kind):
    """The multi-wave 16-bit-output projections go through gemm_ps.cu (one CTA per SM walking the tile list, accumulator
    double-buffered in tensor memory): every tile of every wave must be right, including the ragged last M tile, and the result must
    be bit-identical to the one-tile-per-CTA kernel (same accumulation order over K), which the same product split into row blocks
    of fewer than 300 tiles selects."""
    ops = _ops()
    a = rnd(M, K, seed=1, dtype=_dt())
    w = rnd(N, K, seed=2, std=K ** -0.5, dtype=_dt())
    b = rnd(N, seed=3, std=0.1)
    y = a.float() @ w.float().t() + b
    if kind == "geglu":
        x, gate = y.chunk(2, dim=-1)
        ref = x * torch.nn.functional.gelu(gate)
        w_il, b_il = ops.geglu_interleave(w, b)
        out = ops.geglu_linear(a, w_il, b_il)
        close(out, ref, 1e-2, "persistent geglu")
        # same product through the one-tile kernel: split the rows into launches of fewer than 300 tiles
        step = 128 * max(1, 299 // (N // ops.tile_n(N, ops.EPI_GEGLU)))
        parts = [ops.geglu_linear(a[i:i + step], w_il, b_il) for i in range(0, M, step)]
        assert torch.equal(out, torch.cat(parts)), "persistent and one-tile GEGLU kernels must agree bit for bit"
    else:
        ref = torch.nn.functional.gelu(y) if kind == "gelu" else y
        out = ops.linear(a, w, bias=b, out_dtype=_dt(), gelu=(kind == "gelu"))
        close(out, ref, 1e-2, "persistent linear")
        nobias = ops.linear(a, w, out_dtype=_dt())
        close(nobias, a.float() @ w.float().t(), 1e-2, "persistent linear, no bias")
        step = 128 * max(1, 299 // ((N + ops.tile_n(N) - 1) // ops.tile_n(N)))
        parts = [ops.linear(a[i:i + step], w, bias=b, out_dtype=_dt(), gelu=(kind == "gelu")) for i in range(0, M, step)]
        assert torch.equal(out, torch.cat(parts)), "persistent and one-tile kernels must agree bit for bit"


@pytest.mark.parametrize("C", [320, 512, 1280])
def test_geglu(C):
    ops = _ops()
    M = 384
    a = rnd(M, C, seed=1, dtype=_dt())
    w = rnd(8 * C, C, seed=2, std=C ** -0.5, dtype=_dt())
    b = rnd(8 * C, seed=3, std=0.1)
    y = a.float() @ w.float().t() + b
    x, gate = y.chunk(2, dim=-1)
    ref = x * torch.nn.functional.gelu(gate)
    w_il, b_il = ops.geglu_interleave(w, b)
    out = ops.geglu_linear(a, w_il, b_il)
    assert out.shape == (M, 4 * C)
    close(out, ref, 1e-2, "geglu")


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(16, 32, 32, 64, 320), (16, 16, 16, 640, 640), (16, 8, 8, 1280, 1280), (16, 4, 4, 2560, 1280),
                                             (2, 32, 32, 320, 4), (16, 2, 2, 256, 256), (32, 16, 16, 128, 128)])
def test_conv3x3(NB, H, W, Cin, Cout):
    ops = _ops()
    x = rnd(NB, Cin, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, seed=2, std=(9 * Cin) ** -0.5)
    b = rnd(Cout, seed=3)
    xb, wb = x.to(_dt()), w.to(_dt())
    ref = torch.nn.functional.conv2d(xb.float(), wb.float(), b, padding=1)
    a = xb.permute(0, 2, 3, 1).reshape(NB * H * W, Cin).contiguous()
    wk = wb.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    rb = rnd(NB // 2, Cout, seed=4)
    res = rnd(NB * H * W, Cout, seed=5)
    out = ops.conv3x3(a, wk, NB, H, W, bias=b, rowbias=rb, rows_per_group=2 * H * W, residual=res)
    ref_cl = ref.permute(0, 2, 3, 1).reshape(NB * H * W, Cout) + rb.repeat_interleave(2 * H * W, dim=0) + res
    close(out, ref_cl, 2e-3, "conv3x3")


@pytest.mark.parametrize("B,T,HW,C", [(1, 16, 1024, 320), (2, 16, 256, 640), (1, 16, 64, 1280), (1, 16, 16, 1280), (1, 16, 4, 256)])
def test_conv_t3(B, T, HW, C):
    ops = _ops()
    x = rnd(B, C, T, HW, 1, seed=1)
    w = rnd(C, C, 3, 1, 1, seed=2, std=(3 * C) ** -0.5)
    b = rnd(C, seed=3)
    xb, wb = x.to(_dt()), w.to(_dt())
    ref = torch.nn.functional.conv3d(xb.float(), wb.float(), b, padding=(1, 0, 0))          # [B, C, T, HW, 1]
    a = xb.squeeze(-1).permute(0, 2, 3, 1).reshape(B * T * HW, C).contiguous()
    wk = wb.reshape(C, C, 3).permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    res = rnd(B * T * HW, C, seed=4)
    out = ops.conv_t3(a, wk, B, T, HW, bias=b, residual=res)
    ref_cl = ref.squeeze(-1).permute(0, 2, 3, 1).reshape(B * T * HW, C) + res
    close(out, ref_cl, 2e-3, "conv_t3")


def test_skinny_linear_and_timestep_embedding():
    ops = _ops()
    import sys
    t = torch.tensor([999, 599, 39, 0], device=DEV)
    emb = ops.timestep_embedding(t, 320)
    half = 160
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = t.cpu()[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(DEV)
    assert (emb - ref).abs().max().item() < 2e-4      # fp32 sin/cos of arguments up to 999 rad
    x = rnd(4, 320, seed=1)
    w = rnd(1280, 320, seed=2, std=320 ** -0.5, dtype=_dt())
    b = rnd(1280, seed=3)
    close(ops.skinny_linear(x, w, b, False), x @ w.float().t() + b, 1e-5, "skinny")
    close(ops.skinny_linear(x, w, b, True), torch.nn.functional.silu(x) @ w.float().t() + b, 1e-5, "skinny silu")


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("ns,rows,C,silu", [(16, 1024, 320, True), (1, 16384, 320, True), (16, 64, 1280, False), (16, 16, 2560, True),
                                            (16, 256, 960, True), (2, 4096, 64, True), (16, 64, 1920, False),
                                            # round 2: every GroupNorm shape of the UNet step (per-frame and per-sample statistics), ragged rows, wide maps
                                            (1, 4096, 640, True), (1, 1024, 1280, False), (1, 256, 1280, True), (16, 16, 1280, True),
                                            (16, 1024, 960, True), (16, 1024, 640, True), (16, 256, 1920, True), (4, 16384, 320, False),
                                            (3, 1000, 320, True), (1, 1021, 640, False), (16, 4096, 512, True), (16, 65536, 128, True),
                                            (2, 300, 96, True), (2, 64, 4096, False), (1, 50, 3200, True)])
def test_groupnorm(ns, rows, C, silu):
    """C // 32 in (10, 30): float4 columns straddle two groups; ns = 1 rows: the per-sample statistics of the temporal blocks."""
    ops = _ops()
    x = rnd(ns * rows, C, seed=1) * 3 + 0.5
    g = rnd(C, seed=2) * 0.1 + 1
    b = rnd(C, seed=3) * 0.1
    ref = torch.nn.functional.group_norm(x.view(ns, rows, C).permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1).reshape(ns * rows, C)
    if silu:
        ref = torch.nn.functional.silu(ref)
    out = ops.groupnorm(x, g, b, ns, rows, 1e-5, silu)
    close(out, ref, 6e-3, "groupnorm")


@pytest.mark.parametrize("rows,C", [(16384, 320), (1000, 640), (256, 1280), (512, 512), (64, 64)])
def test_layernorm(rows, C):
    ops = _ops()
    x = rnd(rows, C, seed=1) * 2 + 0.3
    g = rnd(C, seed=2) * 0.1 + 1
    b = rnd(C, seed=3) * 0.1
    add = rnd(rows, C, seed=4) * 0.1
    ref = torch.nn.functional.layer_norm(x, (C,), g, b, 1e-5)
    out = ops.layernorm(x, g, b)
    close(out, ref, 6e-3, "layernorm")
    o1, o2 = ops.layernorm(x, g, b, add=add)
    close(o1, ref, 6e-3, "layernorm.1")
    close(o2, ref + add, 6e-3, "layernorm.2")
    # the second output written into a column block of a wider buffer (the temporal block's K-concatenated operand)
    wide = torch.full((rows, 3 * C), 7.0, device=DEV, dtype=_dt())
    o1s, o2s = ops.layernorm(x, g, b, add=add, out2=wide[:, C:2 * C])
    assert o2s.data_ptr() == wide[:, C:2 * C].data_ptr()
    assert torch.equal(o1s, o1) and torch.equal(wide[:, C:2 * C], o2), "strided out2 must be bit-identical"
    assert (wide[:, :C] == 7).all() and (wide[:, 2 * C:] == 7).all(), "strided out2 wrote outside its column block"


# ------------------------------------------------------------------------------------------------ attention
def ref_attention(q, k, v, heads, mask=None):
    bq, lq, _ = q.shape
    qh = q.float().view(bq, lq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(k.shape[0], -1, heads, 64).permute(0, 2, 1, 3)
    vh = v.float().view(v.shape[0], -1, heads, 64).permute(0, 2, 1, 3)
    sim = qh @ kh.transpose(-1, -2) * 0.125
    if mask is not None:
        sim = sim.masked_fill(~mask[:, None], float("-inf"))
    return (sim.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(bq, lq, heads * 64)


@pytest.mark.parametrize("bq,lq,lk,heads,kv_div", [(16, 1024, 1024, 5, 1), (16, 256, 256, 10, 1), (16, 64, 64, 20, 1), (16, 16, 16, 20, 1),
                                                   (16, 1024, 77, 5, 16), (16, 256, 768, 10, 16), (32, 64, 16, 20, 1), (2, 200, 333, 3, 1)])
def test_attention_dense(bq, lq, lk, heads, kv_div):
    ops = _ops()
    C = heads * 64
    bk = bq // kv_div
    qkv = rnd(bq * lq, 3 * C, seed=1, dtype=_dt())
    q = qkv[:, :C]
    if lq == lk and kv_div == 1:
        k, v = qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        kv = rnd(bk * lk, 2 * C, seed=2, dtype=_dt())
        k, v = kv[:, :C], kv[:, C:]
    out = ops.attention(q, k, v, bq, lq, lk, heads, kv_div=kv_div)
    kk = k.reshape(bk, lk, C).repeat_interleave(kv_div, dim=0)
    vv = v.reshape(bk, lk, C).repeat_interleave(kv_div, dim=0)
    ref = ref_attention(q.reshape(bq, lq, C), kk, vv, heads).reshape(bq * lq, C)
    close(out, ref, 1.5e-2, "attention")
    # accumulate: out += s * attention (image cross-attention sum)
    out2 = ops.attention(q, k, v, bq, lq, lk, heads, kv_div=kv_div, out=out.clone(), out_scale=1.37, accumulate=True)
    close(out2, ref * 2.37, 2e-2, "attention accumulate")


@pytest.mark.parametrize("T,H,W,d,heads,kind", [(16, 8, 8, 32, 4, "pan_yaw"), (16, 4, 4, 64, 20, "orbit"), (16, 16, 16, 16, 2, "dolly"),
                                                (16, 32, 32, 8, 1, "roll_pan_up")])
def test_attention_epipolar(T, H, W, d, heads, kind):
    """Mask evaluated in-kernel from F == attention with the oracle's materialised mask (+ register tokens)."""
    import oracle
    from oracle import camera_oracle
    from camc2v_b200 import synth
    ops = _ops()
    K, w2c = synth.synth_camera(kind, T=T)
    torch.manual_seed(123)
    rel = camera_oracle.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))
    Fm = camera_oracle.fundamental_matrices(K, rel)
    mask = oracle.epipolar_mask(Fm, H, W, d).to(DEV)              # [1, L, L]
    L, C, R = T * H * W, heads * 64, 4
    qkv = rnd(L, 3 * C, seed=1, dtype=_dt())
    reg = rnd(R, 2 * C, seed=2, dtype=_dt())
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm.to(DEV).contiguous(), epi_grid=(T, H, W), epi_d=d)
    kk = torch.cat([reg[:, :C], k], 0)[None]
    vv = torch.cat([reg[:, C:], v], 0)[None]
    mm = torch.nn.functional.pad(mask, (R, 0), value=True)
    ref = ref_attention(q[None], kk, vv, heads, mm)[0]
    close(out, ref, 1.5e-2, "epipolar attention (F)")
    # the same through the reference-format materialised mask
    out_m = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], mask=mask.contiguous())
    close(out_m, ref, 1.5e-2, "epipolar attention (mask)")
    assert torch.equal(out, out_m), "F-evaluated and mask-driven attention must take identical decisions"
    # tile map: skipping whole key tiles must not change a single bit, and the map must cover every unmasked pair
    tmap = ops.epipolar_tile_map(Fm.to(DEV).contiguous(), T, H, W, d)
    if tmap is not None:
        out_t = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm.to(DEV).contiguous(), epi_grid=(T, H, W),
                              epi_d=d, epi_tile_map=tmap)
        assert torch.equal(out, out_t)
        nt, nk = (L + 127) // 128, (L + 63) // 64          # 128-query x 64-key tiles
        bits = torch.tensor([[(int(tmap[0, qt, j >> 5]) >> (j & 31)) & 1 for j in range(nk)] for qt in range(nt)], dtype=torch.bool)
        mpad = torch.zeros(nt * 128, nk * 64, dtype=torch.bool)
        mpad[:L, :L] = mask[0].cpu()
        occ = mpad.view(nt, 128, nk, 64).any(dim=3).any(dim=1)
        assert bool((bits | ~occ).all()), "tile map cleared a tile that contains an attended pair"
        print(f"tile map {kind} {H}x{W}: {bits.float().mean():.3f} of tiles visited, {occ.float().mean():.3f} truly occupied")
        # packed mask: every bit equals the reference's mask, and attention driven by it is bit-identical to the in-kernel predicate
        bm = ops.epipolar_bitmask(Fm.to(DEV).contiguous(), T, H, W, d)                  # [1, nt, L/32, 128] int32
        assert bm is not None and bm.shape == (1, nt, L // 32, 128)
        w = bm[0].cpu().numpy().astype(np.uint32)                                        # [qt, chunk, r]
        unpacked = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)   # [qt, chunk, r, i]
        unpacked = np.transpose(unpacked, (0, 2, 1, 3)).reshape(nt * 128, L)[:L]
        assert np.array_equal(unpacked, mask[0].cpu().numpy()), "packed epipolar mask differs from the reference mask"
        out_b = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm.to(DEV).contiguous(), epi_grid=(T, H, W),
                              epi_d=d, epi_tile_map=tmap, epi_bitmask=bm)
        assert torch.equal(out, out_b)
        out_b2 = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm.to(DEV).contiguous(), epi_grid=(T, H, W),
                               epi_d=d, epi_bitmask=bm)
        assert torch.equal(out, out_b2)
    elif L % 128 == 0:
        # grids without a specialised builder (the 4x4 level, d = 64: a 32-key chunk spans two frames) use the generic packed-mask
        # builder; the attention kernel then never evaluates a predicate
        nt = L // 128
        bm = ops.epipolar_bitmask(Fm.to(DEV).contiguous(), T, H, W, d)
        assert bm is not None and bm.shape == (1, nt, L // 32, 128)
        w = bm[0].cpu().numpy().astype(np.uint32)
        unpacked = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
        unpacked = np.transpose(unpacked, (0, 2, 1, 3)).reshape(nt * 128, L)
        assert np.array_equal(unpacked, mask[0].cpu().numpy()), "generic packed epipolar mask differs from the reference mask"
        out_b = ops.attention(q, k, v, 1, L, L, heads, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm.to(DEV).contiguous(), epi_grid=(T, H, W),
                              epi_d=d, epi_bitmask=bm)
        assert torch.equal(out, out_b)


@pytest.mark.parametrize("B,T,HW,heads", [(1, 16, 1024, 5), (2, 16, 64, 20), (1, 16, 256, 8), (1, 8, 16, 4)])
def test_attention_temporal(B, T, HW, heads):
    ops = _ops()
    C = heads * 64
    qkv = rnd(B * T * HW, 3 * C, seed=1, dtype=_dt())
    out = ops.attention_temporal(qkv, B, T, HW, heads)
    x = qkv.view(B, T, HW, 3, C).permute(3, 0, 2, 1, 4).reshape(3, B * HW, T, C)
    ref = ref_attention(x[0], x[1], x[2], heads).view(B, HW, T, C).permute(0, 2, 1, 3).reshape(B * T * HW, C)
    close(out, ref, 1e-2, "temporal attention")
    wide = torch.full((B * T * HW, 3 * C), 7.0, device=DEV, dtype=_dt())
    ops.attention_temporal(qkv, B, T, HW, heads, out=wide[:, 2 * C:])
    assert torch.equal(wide[:, 2 * C:], out), "strided output must be bit-identical"
    assert (wide[:, :2 * C] == 7).all(), "strided output wrote outside its column block"


# ------------------------------------------------------------------------------------------------ camera (bit-exact)
@pytest.mark.parametrize("kind", ["pan_yaw", "stationary", "dolly", "yaw", "roll_pan_up", "orbit"])
def test_epipolar_mask_bit_exact(kind):
    import oracle
    from oracle import camera_oracle
    from camc2v_b200 import synth
    ops = _ops()
    K, w2c = synth.synth_camera(kind, T=16)
    torch.manual_seed(123)
    rel = camera_oracle.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))
    Fm = camera_oracle.fundamental_matrices(K, rel)
    gold = np.load("tests/golden/masks.npz")
    assert np.array_equal(gold[f"{kind}.F"], Fm.numpy()), "F differs from the reference's (CPU torch build mismatch?)"
    for d in (64, 32, 16, 8):
        hw = 256 // d
        got = ops.epipolar_mask(Fm.to(DEV), hw, hw, d).cpu()
        if d >= 16:
            want = oracle.epipolar_mask(Fm, hw, hw, d)
            assert torch.equal(got, want), f"{kind} d={d}: {(got != want).sum().item()} mismatching mask bits"
        packed = np.packbits(got.numpy(), axis=-1)
        import hashlib
        assert hashlib.sha256(packed.tobytes()).digest() == gold[f"{kind}.d{d}.sha256"].tobytes(), f"{kind} d={d}: sha256 mismatch vs reference"


def test_plucker():
    import oracle
    from camc2v_b200 import synth
    ops = _ops()
    gold = np.load("tests/golden/masks.npz")
    for kind in ("pan_yaw", "orbit"):
        K, _ = synth.synth_camera(kind, T=16)
        rel = torch.from_numpy(gold[f"{kind}.rel_c2w"])
        for mode, key in (("plucker", "plucker_sub"), ("ray", "ray_sub")):
            got = ops.plucker(K.to(DEV), rel.to(DEV), 256, 256, mode).cpu()
            assert (got[..., 3::8, 3::8] - torch.from_numpy(gold[f"{kind}.{key}"])).abs().max().item() < 2e-6
            assert (got - oracle.plucker(K, rel, 256, 256, mode)).abs().max().item() < 2e-6


# ------------------------------------------------------------------------------------------------ glue
def test_layout_and_glue():
    ops = _ops()
    B, Cc, T, H, W = 2, 8, 16, 32, 32
    x = rnd(B, Cc, T, H, W, seed=1)
    cl = ops.to_channels_last(x, B, Cc, T * H * W, Cpad=64, dtype=_dt())
    ref = torch.zeros(B * T * H * W, 64, device=DEV)
    ref[:, :Cc] = x.permute(0, 2, 3, 4, 1).reshape(-1, Cc)
    assert torch.equal(cl.float(), ref.to(_dt()).float())
    y = rnd(B * T * H * W, 4, seed=2)
    back = ops.from_channels_last(y, B, 4, T * H * W)
    assert torch.equal(back.view(B, 4, T, H, W), y.view(B, T, H, W, 4).permute(0, 4, 1, 2, 3))
    a, b = rnd(1000, 320, seed=3), rnd(1000, 640, seed=4)
    of, ob = ops.concat_channels(a, b, True, True)
    assert torch.equal(of, torch.cat([a, b], 1)) and torch.equal(ob, torch.cat([a, b], 1).to(_dt()))
    assert torch.equal(ops.cast_bf16(a), a.to(_dt()))
    img = rnd(4 * 8 * 8, 64, seed=5)
    up = ops.upsample2x(img, 4, 8, 8)
    ref_up = torch.nn.functional.interpolate(img.view(4, 8, 8, 64).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert torch.equal(up.view(4, 16, 16, 64), ref_up.permute(0, 2, 3, 1).to(_dt()))
    col = ops.im2col_s2(img, 4, 8, 8)
    unf = torch.nn.functional.unfold(img.view(4, 8, 8, 64).permute(0, 3, 1, 2), 3, padding=1, stride=2)   # [4, 64*9, 16]
    ref_col = unf.view(4, 64, 9, 16).permute(0, 3, 2, 1).reshape(4 * 16, 9 * 64)
    assert torch.equal(col, ref_col.to(_dt()))


def test_downsample_conv_via_im2col():
    ops = _ops()
    NB, H, W, C = 16, 16, 16, 128
    x = rnd(NB, C, H, W, seed=1).to(_dt()).float()
    w = rnd(C, C, 3, 3, seed=2, std=(9 * C) ** -0.5).to(_dt())
    b = rnd(C, seed=3)
    ref = torch.nn.functional.conv2d(x, w.float(), b, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    col = ops.im2col_s2(x.permute(0, 2, 3, 1).reshape(-1, C).contiguous(), NB, H, W)
    out = ops.linear(col, w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous(), bias=b)
    close(out, ref, 2e-3, "downsample conv")


@pytest.mark.parametrize("shape", [(2, 4, 16, 32, 32), (3, 4, 16, 16, 16), (1, 4, 16, 40, 64), (2, 4, 3, 5, 7)])
@pytest.mark.parametrize("phi", [0.0, 0.7])
def test_cfg_ddim_update(phi, shape):
    """(2|3, ...) samples of up to 98 304 elements take the 8-CTA cluster kernel (per-sample statistics through distributed shared
    memory); the 40x64 latent and the odd-sized one take the one-CTA-per-sample kernel."""
    from oracle import ddim_oracle
    ops = _ops()
    x, ec, eu, nz = (rnd(*shape, seed=s) for s in (1, 2, 3, 4))
    sch = ddim_oracle.ddim_schedule()
    i = 14
    args = (float(sch["alphas"][i]), float(sch["alphas_prev"][i]), float(sch["sigmas"][i]), float(sch["sqrt_one_minus_alphas"][i]))
    xp, p0 = ops.cfg_ddim_update(x, ec, eu, nz, 3.5, phi, *args)
    rxp, rp0 = ddim_oracle.cfg_ddim_update(x.cpu(), ec.cpu(), eu.cpu(), nz.cpu(), args[0], args[1], args[2], args[3], 3.5, phi)
    assert (xp.cpu() - rxp).abs().max().item() < 2e-5 * rxp.abs().max().item()
    assert (p0.cpu() - rp0).abs().max().item() < 2e-5 * rp0.abs().max().item()
