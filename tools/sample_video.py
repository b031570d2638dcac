"""One whole CamContextI2V sample on the B200 kernels, stage by stage (synthetic weights and inputs, shipped sizes):

    Resampler (image tokens of the reference + 2 context frames)  ->  conditional epipolar mask + MultiLatentEpipolarAdaptor (c_concat)
    ->  camera condition (F, tile maps, packed masks)  ->  25-step DDIM sampling with CFG (CUDA graph)  ->  VAE decode to 16 RGB frames

    python tools/sample_video.py [--steps 25]
Prints one JSON line with the time of every stage (CUDA events / host clock around synchronised stages).  The CLIP encoders, the
VAE *encoding* of the input frames and the CameraPoseEncoder are outside (their outputs are synthetic tensors of the right shape).
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
    pluker = [synth.synth_tensor(f"sample.pl{i}", (1, c, 16, 32 >> i, 32 >> i), 4, std=0.1).to(dev) for i, c in enumerate((320, 640, 1280, 1280))]
    x_T = synth.synth_tensor("sample.xT", (1, 4, 16, 32, 32), 5).to(dev)
    fs = torch.full((1,), 3, dtype=torch.long, device=dev)
    cond_idx = torch.zeros(1, dtype=torch.long)
    stages = {}

    def warm():          # builds the weight packs of every module once (not part of a sample)
        resampler(clip_img)
        m = ops.epipolar_mask(camera.conditional_fundamental_matrices(K, w2c, w2c_cond, cond_idx).to(dev), 32, 32, 8)
        adaptor(z_cond, m)
        vae.decode(x_T[0].permute(1, 0, 2, 3).contiguous())
    warm()

    img_tok, stages["resampler_ms"] = timed(lambda: resampler(clip_img))              # [3, 256, 1024]

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
