"""One whole CamContextI2V sample on the B200 kernels, stage by stage (synthetic weights and inputs, shipped sizes):

    Resampler (image tokens of the reference + 2 context frames)  ->  conditional epipolar mask + MultiLatentEpipolarAdaptor (c_concat)
    ->  Pluecker embedding + CameraPoseEncoder (pluker_embedding_features)  ->  camera condition (F, tile maps, packed masks)  ->  25-step DDIM sampling with CFG (CUDA graph)  ->  VAE decode to 16 RGB frames

    python tools/sample_video.py [--steps 25]
Prints one JSON line with the time of every stage (CUDA events / host clock around synchronised stages).  The CLIP encoders, the
VAE *encoding* of the input frames are outside (their outputs are synthetic tensors of the right shape).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from camc2v_b200 import camera, ops, synth  # noqa: E402
from camc2v_b200.adaptor import MultiLatentEpipolarAdaptor  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402
from camc2v_b200.modules import build_unet  # noqa: E402
from camc2v_b200.pose_encoder import CameraPoseEncoder  # noqa: E402
from camc2v_b200.resampler import Resampler  # noqa: E402
from camc2v_b200.sampler import DDIMSampler, DenoiserModel  # noqa: E402
from camc2v_b200.vae import AutoencoderKLDecoder  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=25)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = UNetConfig()
    unet = build_unet(cfg)
    synth.fill_module_(unet, seed=0)
    model = DenoiserModel(unet).to(dev)
    sampler = DDIMSampler(model)
    adaptor = MultiLatentEpipolarAdaptor(query_dim=512, num_queries=1024, video_length=16, embedding_dim=4, output_dim=4, depth=12,
                                         timestep_embedding_type="sinusoidal_embedded")
    synth.fill_module_(adaptor, seed=5)
    adaptor = adaptor.to(dev)
    resampler = Resampler(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024, ff_mult=4,
                          video_length=16, use_timestep_emb=True)
    synth.fill_module_(resampler, seed=6)
    resampler = resampler.to(dev)
    vae = AutoencoderKLDecoder(dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                                    num_res_blocks=2, attn_resolutions=[], dropout=0.0))
    synth.fill_module_(vae, seed=7)
    vae = vae.to(dev)

    # synthetic per-sample inputs (what the frozen CLIP / VAE-encode / pose-encoder stages would deliver)
    K, w2c = synth.synth_camera("pan_yaw", T=16, B=1)
    _, w2c_o = synth.synth_camera("orbit", T=16, B=1)
    w2c_cond = w2c_o[:, [5, 11]].contiguous()
    clip_img = synth.synth_tensor("sample.clip", (3, 257, 1280), 1).to(dev)          # reference frame + 2 context frames
    text = synth.synth_tensor("sample.text", (1, 77, 1024), 2).to(dev)
    z_cond = synth.synth_tensor("sample.zcond", (1, 3 * 1024, 4), 3).to(dev)           # their VAE latents as tokens
    pose_enc = CameraPoseEncoder(downscale_factor=8, channels=[320, 640, 1280, 1280], nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False,
                                 compression_factor=1, temporal_attention_nhead=8, attention_block_types=["Temporal_Self"],
                                 temporal_position_encoding=True, temporal_position_encoding_max_len=16)
    pe = {k: v.clone() for k, v in pose_enc.state_dict().items() if k.endswith("pos_encoder.pe")}
    synth.fill_module_(pose_enc, seed=8)
    pose_enc.load_state_dict(pe, strict=False)
    pose_enc = pose_enc.to(dev)
    x_T = synth.synth_tensor("sample.xT", (1, 4, 16, 32, 32), 5).to(dev)
    fs = torch.full((1,), 3, dtype=torch.long, device=dev)
    cond_idx = torch.zeros(1, dtype=torch.long)
    stages = {}

    def run_pose():      # camcontexti2v.py:556-561: ray_condition -> pose_encoder -> '(b f) c h w -> b c f h w'
        pl = ops.plucker(K.to(dev), camera.relative_c2w(w2c, cond_idx).to(dev), 256, 256)
        return [f.view(1, 16, f.shape[1], f.shape[2], f.shape[3]).permute(0, 2, 1, 3, 4).contiguous() for f in pose_enc(pl)]

    def warm():          # builds the weight packs of every module once (not part of a sample)
        resampler(clip_img)
        m = ops.epipolar_mask(camera.conditional_fundamental_matrices(K, w2c, w2c_cond, cond_idx).to(dev), 32, 32, 8)
        adaptor(z_cond, m)
        vae.decode(x_T[0].permute(1, 0, 2, 3).contiguous())
        run_pose()

    warm()

    img_tok, stages["resampler_ms"] = timed(lambda: resampler(clip_img))              # [3, 256, 1024]
    pluker, stages["plucker_and_pose_encoder_ms"] = timed(run_pose)

    def run_adaptor():
        Fm = camera.conditional_fundamental_matrices(K, w2c, w2c_cond, cond_idx).to(dev)
        mask = ops.epipolar_mask(Fm, 32, 32, 8)
        y = adaptor(z_cond, mask)                                                       # [1, 16*1024, 4]
        return y.view(1, 16, 32, 32, 4).permute(0, 4, 1, 2, 3).contiguous()
    c_concat, stages["cond_mask_and_adaptor_ms"] = timed(run_adaptor)

    def run_camera():
        return camera.camera_condition(K, w2c, cond_idx, 256, 256, pluker_embedding_features=pluker, device=dev)
    cam, stages["camera_condition_ms"] = timed(run_camera)
    ctx_cond = torch.cat([text, img_tok.reshape(1, 3 * 256, 1024)], dim=1)              # 77 + 256 * (1 + 2) tokens
    ctx_unc = torch.cat([torch.zeros_like(text), img_tok[:1].reshape(1, 256, 1024)], dim=1)
    cond = {"c_crossattn": [ctx_cond], "c_concat": [c_concat], "camera_condition": cam}
    uc = {"c_crossattn": [ctx_unc], "c_concat": [c_concat]}

    def run_loop():
        torch.manual_seed(20230211)
        return sampler.sample(args.steps, 1, (4, 16, 32, 32), conditioning=cond, eta=1.0, verbose=False, x_T=x_T, unconditional_guidance_scale=3.5,
                              unconditional_conditioning=uc, fs=fs, timestep_spacing="uniform_trailing", guidance_rescale=0.7,
                              enable_camera_condition=True, use_cuda_graph=True)[0]
    _, stages["first_loop_incl_graph_capture_ms"] = timed(run_loop)
    z0, stages["ddim_loop_ms"] = timed(run_loop)
    z0b, _ = timed(run_loop)
    zf = z0[0].permute(1, 0, 2, 3).contiguous()
    _, stages["first_vae_decode_after_loop_ms"] = timed(lambda: vae.decode(zf))       # includes cudaMalloc of its 0.5 GB activations
    frames, stages["vae_decode_ms"] = timed(lambda: vae.decode(zf))                     # [16, 3, 256, 256], allocator warm
    per_sample = sum(v for k, v in stages.items() if not k.startswith("first_"))
    line = {"metric": "sample_ms", "value": per_sample, "unit": f"ms per 16-frame 256x256 video ({args.steps} DDIM steps, CFG 3.5), one B200",
            "videos_per_s": 1e3 / per_sample, "stages": stages, "deterministic": bool(torch.equal(z0, z0b)),
            "finite": bool(torch.isfinite(frames).all()), "frames_shape": list(frames.shape), "dtype": ops._lib.OPERANDS}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
