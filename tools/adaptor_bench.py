"""Once-per-sample MultiLatentEpipolarAdaptor at the shipped size (camcontexti2v_256.yaml:141-152:
query_dim 512, depth 12, 16 x 1024 queries, 1 reference + 2 context frames = 3072 context tokens) on one B200, plus the
conditional mask build; the CPU oracle (port of the reference) is timed on ONE layer-pair-sized sample and scaled by depth.

    python tools/adaptor_bench.py [--no-cpu]
Prints one JSON line.  CUDA events on the launching stream, 2 warm-up + 5 timed forwards.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from camc2v_b200 import camera, ops, synth  # noqa: E402
from camc2v_b200.adaptor import MultiLatentEpipolarAdaptor  # noqa: E402

KW = dict(query_dim=512, num_queries=1024, video_length=16, embedding_dim=4, output_dim=4, depth=12, checkpoint=True,
          timestep_embedding_type="sinusoidal_embedded", use_plucker_embedding=False)


def flops(depth, Lq, Lk, D, inner=512, ff=4):
    per = 2 * Lq * D * inner + 2 * (Lk + 2) * D * 2 * inner + 4 * Lq * (Lk + 2) * inner + 2 * Lq * inner * D + 2 * 2 * Lq * D * ff * D
    return depth * per


def main():
    dev = torch.device("cuda", 0)
    m = MultiLatentEpipolarAdaptor(**KW)
    synth.fill_module_(m, seed=5)
    m = m.to(dev)
    K, w2c = synth.synth_camera("pan_yaw", T=16, B=1)
    _, w2c_o = synth.synth_camera("orbit", T=16, B=1)
    w2c_cond = w2c_o[:, [5, 11]].contiguous()
    z = synth.synth_tensor("adaptor.z", (1, 3 * 1024, 4), 9).to(dev)

    def build_mask():
        Fm = camera.conditional_fundamental_matrices(K, w2c, w2c_cond, torch.zeros(1, dtype=torch.long)).to(dev)
        return ops.epipolar_mask(Fm, 32, 32, 8)

    mask = build_mask()
    for _ in range(2):
        y = m(z, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = m(z, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    t0 = time.perf_counter()
    for _ in range(5):
        build_mask()
    torch.cuda.synchronize()
    ms_mask = (time.perf_counter() - t0) / 5 * 1e3
    fl = flops(12, 16384, 3072, 512)
    line = {"metric": "adaptor_ms_per_sample", "value": ms, "unit": "ms (MultiLatentEpipolarAdaptor forward, 16384 queries x 3072 context tokens, depth 12)",
            "mask_build_ms": ms_mask, "mask_density": float(mask.float().mean()), "algorithmic_tflop": fl / 1e12,
            "tflops": fl / ms / 1e9, "finite": bool(torch.isfinite(y).all()), "dtype": ops._lib.OPERANDS}
    if "--no-cpu" not in sys.argv:
        from oracle import adaptor_oracle
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        zc, mc = z.cpu(), mask.cpu()
        t0 = time.perf_counter()
        adaptor_oracle.adaptor_forward(sd, zc, mc, depth=1)
        s1 = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": s1 * 12 * 1e3, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle (CPU port of the reference, fp32) on 1 of the 12 layers: {s1:.2f} s, scaled x12"}
    print(json.dumps(line))


def resampler_main():
    """Resampler at the shipped size (camcontexti2v_256.yaml:110-122): 257 CLIP tokens of one frame -> 256 context tokens."""
    from camc2v_b200.resampler import Resampler
    dev = torch.device("cuda", 0)
    kw = dict(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024, ff_mult=4, video_length=16,
              use_timestep_emb=True)
    m = Resampler(**kw)
    synth.fill_module_(m, seed=6)
    m = m.to(dev)
    x = synth.synth_tensor("resampler.x", (3, 257, 1280), 10).to(dev)          # reference frame + 2 context frames
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    line = {"metric": "resampler_ms_per_sample", "value": e0.elapsed_time(e1) / 10, "unit": "ms (Resampler forward, 3 frames x 257 CLIP tokens -> 3 x 256 x 1024)",
            "finite": bool(torch.isfinite(y).all()), "dtype": ops._lib.OPERANDS}
    if "--no-cpu" not in sys.argv:
        from oracle import resampler_oracle
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        t0 = time.perf_counter()
        resampler_oracle.resampler_forward(sd, x.cpu(), depth=4, heads=12)
        line["cpu_baseline"] = {"value": (time.perf_counter() - t0) * 1e3, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle (CPU port of the reference, fp32), whole forward"}
    print(json.dumps(line))


def pose_encoder_main():
    """CameraPoseEncoder at the shipped size (camcontexti2v_256.yaml:125-139): Pluecker embedding [1, 6, 16, 256, 256] -> 4 feature maps."""
    from camc2v_b200.pose_encoder import CameraPoseEncoder
    dev = torch.device("cuda", 0)
    kw = dict(downscale_factor=8, channels=[320, 640, 1280, 1280], nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False, compression_factor=1,
              temporal_attention_nhead=8, attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
              temporal_position_encoding_max_len=16)
    m = CameraPoseEncoder(**kw)
    pe = {k: v.clone() for k, v in m.state_dict().items() if k.endswith("pos_encoder.pe")}
    synth.fill_module_(m, seed=8)
    m.load_state_dict(pe, strict=False)
    m = m.to(dev)
    K, w2c = synth.synth_camera("orbit", T=16, H=256, W=256, B=1)
    from camc2v_b200 import camera
    x = ops.plucker(K.to(dev), camera.relative_c2w(w2c, torch.zeros(1, dtype=torch.long)).to(dev), 256, 256)
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    line = {"metric": "pose_encoder_ms_per_sample", "value": e0.elapsed_time(e1) / 10,
            "unit": "ms (CameraPoseEncoder forward, [1,6,16,256,256] -> 320x32x32 / 640x16x16 / 1280x8x8 / 1280x4x4 per frame)",
            "finite": bool(all(torch.isfinite(f).all() for f in y)), "shapes": [list(f.shape) for f in y], "dtype": ops._lib.OPERANDS}
    if "--no-cpu" not in sys.argv:
        from oracle import pose_encoder_oracle
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        t0 = time.perf_counter()
        fo = pose_encoder_oracle.pose_encoder_forward(sd, x.cpu())
        line["cpu_baseline"] = {"value": (time.perf_counter() - t0) * 1e3, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle (CPU port of the reference, fp32), whole forward"}
        line["rel_l2_vs_oracle"] = [float((a.cpu().double() - b.double()).norm() / b.double().norm()) for a, b in zip(y, fo)]
    print(json.dumps(line))


def vae_main():
    """decode_first_stage at the shipped size (camcontexti2v_256.yaml:74-93): 16 latent frames [4, 32, 32] -> 16 x [3, 256, 256]."""
    from camc2v_b200.vae import AutoencoderKLDecoder
    dev = torch.device("cuda", 0)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
              attn_resolutions=[], dropout=0.0)
    m = AutoencoderKLDecoder(dd)
    synth.fill_module_(m, seed=7)
    m = m.to(dev)
    z = synth.synth_tensor("vae.z", (16, 4, 32, 32), 11).to(dev)
    for _ in range(2):
        y = m.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = m.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    # conv FLOPs per frame (2 * HW * 9 * Cin * Cout), the dominant term
    def conv(hw, ci, co, k=9):
        return 2.0 * hw * k * ci * co
    fl = conv(1024, 4, 512) + 2 * 2 * conv(1024, 512, 512) + 3 * 2 * conv(1024, 512, 512) + conv(4096, 512, 512) + 3 * 2 * conv(4096, 512, 512) + \
        conv(16384, 512, 512) + conv(16384, 512, 256) + 5 * conv(16384, 256, 256) + conv(65536, 256, 256) + conv(65536, 256, 128) + \
        5 * conv(65536, 128, 128) + conv(65536, 128, 3) + 4 * 2.0 * 1024 * 512 * 512 + 2 * 2.0 * 1024 * 1024 * 512
    line = {"metric": "vae_decode_ms_per_video", "value": ms, "unit": "ms (AutoencoderKL.decode, 16 frames 32x32 latent -> 256x256 RGB)",
            "algorithmic_tflop": 16 * fl / 1e12, "tflops": 16 * fl / ms / 1e9, "finite": bool(torch.isfinite(y).all()), "dtype": ops._lib.OPERANDS}
    if "--no-cpu" not in sys.argv:
        from oracle import vae_oracle
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        t0 = time.perf_counter()
        yo = vae_oracle.decode(sd, z[:1].cpu(), dd["ch_mult"], dd["num_res_blocks"])
        s1 = time.perf_counter() - t0
        a, b = y[:1].cpu().double(), yo.double()
        line["cpu_baseline"] = {"value": s1 * 16 * 1e3, "unit": "ms", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle (CPU port of the reference, fp32) on 1 of the 16 frames: {s1:.2f} s, scaled x16"}
        line["parity_full_size_frame0"] = {"rel_l2": float((a - b).norm() / b.norm()), "max_norm": float((a - b).abs().max() / b.abs().max())}
    print(json.dumps(line))


if __name__ == "__main__":
    if "--vae" in sys.argv:
        vae_main()
    elif "--resampler" in sys.argv:
        resampler_main()
    elif "--pose-encoder" in sys.argv:
        pose_encoder_main()
    else:
        main()
