"""Per-kernel microbenchmarks on the shapes of one CamContextI2V UNet pass (B=1): time, TFLOP/s, GB/s.

    python tools/kernel_bench.py [gemm] [conv] [norm] [attn]

CUDA events on the launching stream, 3 warm-up + 10 timed launches, L2 flushed between launches.
"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from camc2v_b200 import ops  # noqa: E402

DEV = "cuda"
FLUSH = None


def timeit(fn, n=10):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us (median)


def rb(*shape, dtype=None):
    dtype = dtype or ops.BF16
    return torch.randn(*shape, device=DEV).to(dtype)


def bench_gemm():
    print("== linear: M K N  [out dtype, residual]")
    shapes = [
        (16384, 320, 320, "f32", True), (16384, 960, 320, "f32", True), (16384, 320, 960, "bf16", False), (16384, 1280, 320, "f32", True),
        (4096, 640, 640, "f32", True), (4096, 640, 1920, "bf16", False), (4096, 2560, 640, "f32", True),
        (1024, 1280, 1280, "f32", True), (1024, 1280, 3840, "bf16", False), (1024, 5120, 1280, "f32", True),
        (256, 1280, 1280, "f32", True), (256, 5120, 1280, "f32", True), (768, 1024, 2560, "bf16", False), (77, 1024, 640, "bf16", False),
    ]
    for M, K, N, od, res in shapes:
        a, w, b = rb(M, K), rb(N, K), rb(N, dtype=torch.float32)
        r = rb(M, N, dtype=torch.float32) if res else None
        odt = torch.float32 if od == "f32" else ops.BF16
        us = timeit(lambda: ops.linear(a, w, bias=b, residual=r, out_dtype=odt))
        fl = 2.0 * M * N * K
        by = M * K * 2 + N * K * 2 + M * N * (4 if od == "f32" else 2) + (M * N * 4 if res else 0)
        print(f"  {M:6d} {K:5d} {N:5d} {od:4s} res={int(res)}  {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s")
    print("== geglu: M C")
    for M, C in [(16384, 320), (4096, 640), (1024, 1280), (256, 1280)]:
        a, w, b = rb(M, C), rb(8 * C, C), rb(8 * C, dtype=torch.float32)
        wi, bi = ops.geglu_interleave(w, b)
        us = timeit(lambda: ops.geglu_linear(a, wi, bi))
        print(f"  {M:6d} {C:5d}  {us:8.1f} us  {2.0 * M * 8 * C * C / us / 1e6:7.1f} TFLOP/s")


def bench_conv():
    print("== conv3x3: NB H W Cin Cout")
    for NB, H, W, Ci, Co in [(16, 32, 32, 320, 320), (16, 32, 32, 640, 320), (16, 32, 32, 960, 320), (16, 16, 16, 640, 640), (16, 16, 16, 1280, 640),
                             (16, 16, 16, 1920, 640), (16, 8, 8, 1280, 1280), (16, 8, 8, 2560, 1280), (16, 4, 4, 1280, 1280), (16, 4, 4, 2560, 1280)]:
        a, w, b = rb(NB * H * W, Ci), rb(Co, 9 * Ci), rb(Co, dtype=torch.float32)
        r = rb(NB * H * W, Co, dtype=torch.float32)
        us = timeit(lambda: ops.conv3x3(a, w, NB, H, W, bias=b, residual=r))
        print(f"  {NB:3d} {H:3d} {W:3d} {Ci:5d} {Co:5d}  {us:8.1f} us  {2.0 * NB * H * W * 9 * Ci * Co / us / 1e6:7.1f} TFLOP/s")
    print("== conv_t3: B T HW C")
    for B, T, HW, C in [(1, 16, 1024, 320), (1, 16, 256, 640), (1, 16, 64, 1280), (1, 16, 16, 1280)]:
        a, w, b = rb(B * T * HW, C), rb(C, 3 * C), rb(C, dtype=torch.float32)
        us = timeit(lambda: ops.conv_t3(a, w, B, T, HW, bias=b))
        print(f"  {B:3d} {T:3d} {HW:5d} {C:5d}  {us:8.1f} us  {2.0 * B * T * HW * 3 * C * C / us / 1e6:7.1f} TFLOP/s")


def bench_norm():
    print("== groupnorm+silu: ns rows C  (algorithmic bytes = 4B read + 2B write per element)")
    for ns, rows, C in [(16, 1024, 320), (1, 16384, 320), (16, 1024, 640), (16, 1024, 960), (16, 256, 640), (1, 4096, 640), (16, 256, 1920),
                        (16, 64, 1280), (1, 1024, 1280), (16, 64, 2560), (16, 16, 1280), (16, 16, 2560)]:
        x, g, b = rb(ns * rows, C, dtype=torch.float32), rb(C, dtype=torch.float32), rb(C, dtype=torch.float32)
        us = timeit(lambda: ops.groupnorm(x, g, b, ns, rows, 1e-5, True))
        print(f"  {ns:3d} {rows:6d} {C:5d}  {us:8.1f} us  {ns * rows * C * 6 / us / 1e3:7.1f} GB/s")
    print("== layernorm: rows C")
    for rows, C in [(16384, 320), (4096, 640), (1024, 1280), (256, 1280)]:
        x, g, b = rb(rows, C, dtype=torch.float32), rb(C, dtype=torch.float32), rb(C, dtype=torch.float32)
        us = timeit(lambda: ops.layernorm(x, g, b))
        print(f"  {rows:6d} {C:5d}  {us:8.1f} us  {rows * C * 6 / us / 1e3:7.1f} GB/s")


def bench_attn():
    from camc2v_b200 import camera, synth
    print("== attention: bq lq lk heads kv_div")
    for bq, lq, lk, h, div in [(16, 1024, 1024, 5, 1), (16, 256, 256, 10, 1), (16, 64, 64, 20, 1), (16, 1024, 77, 5, 16), (16, 1024, 768, 5, 16),
                               (16, 256, 768, 10, 16), (16, 64, 768, 20, 16), (16, 1024, 16, 5, 1)]:
        C = h * 64
        q = rb(bq * lq, C)
        kv = rb((bq // div) * lk, 2 * C)
        us = timeit(lambda: ops.attention(q, kv[:, :C], kv[:, C:], bq, lq, lk, h, kv_div=div))
        print(f"  {bq:3d} {lq:5d} {lk:5d} {h:3d} {div:3d}  {us:8.1f} us  {4.0 * bq * lq * lk * C / us / 1e6:7.1f} TFLOP/s")
    print("== epipolar attention: T H W d heads trajectory")
    for kind in ("pan_yaw", "dolly", "stationary"):
        for T, H, d, h in [(16, 32, 8, 5), (16, 16, 16, 10), (16, 8, 32, 20), (16, 4, 64, 20)]:
            L, C = T * H * H, h * 64
            qkv, reg = rb(L, 3 * C), rb(4, 2 * C)
            K, w2c = synth.synth_camera(kind, T=T)
            torch.manual_seed(123)
            Fm = camera.fundamental_matrices(K, camera.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))).to(DEV).contiguous()
            tmap = ops.epipolar_tile_map(Fm, T, H, H, d)       # once per sample, as in the model path
            bmask = ops.epipolar_bitmask(Fm, T, H, H, d)
            us = timeit(lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], 1, L, L, h, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm,
                                              epi_grid=(T, H, H), epi_d=d, epi_tile_map=tmap, epi_bitmask=bmask))
            vis = 1.0 if tmap is None else float(sum(bin(int(w) & 0xffffffff).count("1") for w in tmap[..., :-1].flatten().tolist())) / ((L // 128) * (L // 64))
            print(f"  {kind:11s} {T:3d} {H:3d} {d:3d} {h:3d}  {us:8.1f} us  {4.0 * L * (L + 4) * C / us / 1e6:7.1f} TFLOP/s (dense-equivalent; "
                  f"{vis * 100:.1f}% of tiles visited -> {4.0 * L * (L + 4) * C * vis / us / 1e6:7.1f} executed)")
    print("== temporal attention: B T HW heads")
    for B, T, HW, h in [(1, 16, 1024, 5), (1, 16, 1024, 8), (1, 16, 256, 10), (1, 16, 64, 20)]:
        qkv = rb(B * T * HW, 3 * h * 64)
        us = timeit(lambda: ops.attention_temporal(qkv, B, T, HW, h))
        print(f"  {B:3d} {T:3d} {HW:5d} {h:3d}  {us:8.1f} us  {B * T * HW * h * 64 * 2 * 4 / us / 1e3:7.1f} GB/s")


def single(which):
    """One representative launch series (for `ncu --set full -s 3 -c 1`)."""
    from camc2v_b200 import camera, synth
    if which == "lin0":
        a, w, b, r = rb(16384, 320), rb(320, 320), rb(320, dtype=torch.float32), rb(16384, 320, dtype=torch.float32)
        fn = lambda: ops.linear(a, w, bias=b, residual=r)
    elif which == "lincat0":        # the temporal block's fused output projection: K-concatenated [n + p | attn1 | epipolar], 32x32 level
        a, w, b, r = rb(16384, 960), rb(320, 960), rb(320, dtype=torch.float32), rb(16384, 320, dtype=torch.float32)
        fn = lambda: ops.linear(a, w, bias=b, residual=r)
    elif which == "conv0":
        a, w, b, r = rb(16384, 320), rb(320, 2880), rb(320, dtype=torch.float32), rb(16384, 320, dtype=torch.float32)
        fn = lambda: ops.conv3x3(a, w, 16, 32, 32, bias=b, residual=r)
    elif which == "gn0":
        x, g, b = rb(16384, 320, dtype=torch.float32), rb(320, dtype=torch.float32), rb(320, dtype=torch.float32)
        fn = lambda: ops.groupnorm(x, g, b, 16, 1024, 1e-5, True)
    elif which == "epi0":
        T, H, d, h = 16, 32, 8, 5
        L, C = T * H * H, h * 64
        qkv, reg = rb(L, 3 * C), rb(4, 2 * C)
        K, w2c = synth.synth_camera("pan_yaw", T=T)
        torch.manual_seed(123)
        Fm = camera.fundamental_matrices(K, camera.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))).to(DEV).contiguous()
        fn = lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], 1, L, L, h, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm,
                                   epi_grid=(T, H, H), epi_d=d)
    elif which == "epi0map":
        T, H, d, h = 16, 32, 8, 5
        L, C = T * H * H, h * 64
        qkv, reg = rb(L, 3 * C), rb(4, 2 * C)
        K, w2c = synth.synth_camera("pan_yaw", T=T)
        torch.manual_seed(123)
        Fm = camera.fundamental_matrices(K, camera.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))).to(DEV).contiguous()
        tmap = ops.epipolar_tile_map(Fm, T, H, H, d)
        bmask = ops.epipolar_bitmask(Fm, T, H, H, d)
        fn = lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], 1, L, L, h, k2=reg[:, :C], v2=reg[:, C:], epi_F=Fm,
                                   epi_grid=(T, H, H), epi_d=d, epi_tile_map=tmap, epi_bitmask=bmask)
    elif which == "geglu0":
        a, w, b = rb(16384, 320), rb(2560, 320), rb(2560, dtype=torch.float32)
        wi, bi = ops.geglu_interleave(w, b)
        fn = lambda: ops.geglu_linear(a, wi, bi)
    elif which == "qkv0":
        a, w = rb(16384, 320), rb(960, 320)
        fn = lambda: ops.linear(a, w, out_dtype=ops.BF16)
    elif which == "qkv1":
        a, w = rb(4096, 640), rb(1920, 640)
        fn = lambda: ops.linear(a, w, out_dtype=ops.BF16)
    elif which == "geglu1":
        a, w, b = rb(4096, 640), rb(5120, 640), rb(5120, dtype=torch.float32)
        wi, bi = ops.geglu_interleave(w, b)
        fn = lambda: ops.geglu_linear(a, wi, bi)
    elif which == "attn0":
        q, kv = rb(16 * 1024, 320), rb(16 * 1024, 640)
        fn = lambda: ops.attention(q, kv[:, :320], kv[:, 320:], 16, 1024, 1024, 5)
    elif which in ("attn77", "attn16"):
        lk, div = (77, 16) if which == "attn77" else (16, 1)
        q, kv = rb(16 * 1024, 320), rb((16 // div) * lk, 640)
        fn = lambda: ops.attention(q, kv[:, :320], kv[:, 320:], 16, 1024, lk, 5, kv_div=div)
    else:
        raise SystemExit(which)
    print(which, f"{timeit(fn):.1f} us")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "single":
        single(sys.argv[2])
        sys.exit(0)
    which = sys.argv[1:] or ["gemm", "conv", "norm", "attn"]
    torch.manual_seed(0)
    if "gemm" in which:
        bench_gemm()
    if "conv" in which:
        bench_conv()
    if "norm" in which:
        bench_norm()
    if "attn" in which:
        bench_attn()
