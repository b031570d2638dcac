"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

    python oracle/refgen/make_golden.py [--only masks|unet_small|ddim|unet_full|loop_full] [--check-oracle]

The reference has no tests or golden vectors of its own; these files are outputs of the
reference's own code (imported from /root/reference by ref_harness.py) on seeded synthetic inputs that
the tests regenerate with camc2v_b200.synth.  With --check-oracle every golden is also compared with
oracle/ on the spot and the deviation printed (recorded in DESIGN.md).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh  # noqa: E402
from camc2v_b200 import synth  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRAJ = ["pan_yaw", "stationary", "dolly", "yaw", "roll_pan_up", "orbit"]
SMALL_UNET = dict(model_channels=64)
SMALL_CFG = UNetConfig(model_channels=64, origin_h=128, origin_w=128)
SMALL_HW = 16
PERTURB_SEED = 123


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel_err(a: torch.Tensor, b: torch.Tensor):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


# ---------------------------------------------------------------------------------------------
def ref_camera(model, K, w2c, cond_idx, H, W, resolutions):
    """The geometry half of get_batch_input_camera_condition_process (camcontexti2v.py:525-554), calling
    the reference's own methods in the reference's order."""
    with torch.no_grad():
        c2w = w2c.float().inverse()
        rel = model.get_relative_pose(c2w, cond_idx, mode="left", normalize_T0=False)
        rel[:, :, :3, 3] = rel[:, :, :3, 3] * 1.0
        pairs = model.get_relative_c2w_RT_pairs(rel)
        R = pairs[..., :3, :3]
        t = pairs[..., :3, 3:4]
        t = model.add_small_perturbation(t, epsilon=1e-6)
        F = model.get_fundamental_matrix(K.float().unsqueeze(1), R, t)
        T = w2c.shape[1]
        masks = {int(8 * ds): model.get_epipolar_mask(F, T, H // int(8 * ds), W // int(8 * ds), int(8 * ds)) for ds in resolutions}
        return rel, F, masks


def gen_masks(model, check):
    import oracle
    from oracle import camera_oracle

    out = {}
    for kind in TRAJ:
        K, w2c = synth.synth_camera(kind, T=16, H=256, W=256, B=1)
        cond = torch.zeros(1, dtype=torch.long)
        torch.manual_seed(PERTURB_SEED)
        rel, F, masks = ref_camera(model, K, w2c, cond, 256, 256, (8, 4, 2, 1))
        out[f"{kind}.F"] = F.numpy()
        out[f"{kind}.rel_c2w"] = rel.numpy()
        for d, m in masks.items():
            mn = m.numpy()
            packed = np.packbits(mn, axis=-1)
            out[f"{kind}.d{d}.sha256"] = np.frombuffer(bytes.fromhex(sha(packed)), dtype=np.uint8)
            out[f"{kind}.d{d}.rowsum"] = mn.sum(-1).astype(np.int32)
            out[f"{kind}.d{d}.colsum"] = mn.sum(-2).astype(np.int32)
            if d >= 32:
                out[f"{kind}.d{d}.packed"] = packed
            print(f"  mask {kind} d={d} shape={tuple(mn.shape)} density={mn.mean():.4f}")
        pl = model.ray_condition(K.float(), rel, 256, 256, "cpu")
        out[f"{kind}.plucker_sub"] = pl[..., 3::8, 3::8].numpy()
        model.camera_embedding = "ray"
        ry = model.ray_condition(K.float(), rel, 256, 256, "cpu")
        model.camera_embedding = "plucker"
        out[f"{kind}.ray_sub"] = ry[..., 3::8, 3::8].numpy()
        if check:
            torch.manual_seed(PERTURB_SEED)
            F2, m2, rel2 = camera_oracle.camera_condition_masks(K, w2c, cond)
            print(f"    oracle: F bit-equal={torch.equal(F2, F)} rel bit-equal={torch.equal(rel2, rel)}", end="")
            for d in masks:
                print(f" d{d}:mismatch={(m2[d] != masks[d]).sum().item()}", end="")
            p2 = oracle.plucker(K, rel, 256, 256)
            print(f" plucker maxabs={float((p2 - pl).abs().max()):.2e}")
    np.savez_compressed(os.path.join(GOLD, "masks.npz"), **out)


# ---------------------------------------------------------------------------------------------
from camc2v_b200.testing import synth_unet_inputs as synth_inputs  # noqa: E402


def build_ref(unet_overrides, origin):
    model = rh.build_reference_model(unet_overrides=unet_overrides)
    if origin is not None:
        for m in model.model.diffusion_model.modules():
            if m.__class__.__name__ == "Epipolar":
                m.origin_h = m.origin_w = origin
    synth.fill_module_(model.model.diffusion_model, seed=0)
    return model


def camera_condition(model, hw_img, kind, pluker):
    K, w2c = synth.synth_camera(kind, T=16, H=hw_img, W=hw_img, B=1)
    torch.manual_seed(PERTURB_SEED)
    rel, F, masks = ref_camera(model, K, w2c, torch.zeros(1, dtype=torch.long), hw_img, hw_img, (8, 4, 2, 1))
    return {"pluker_embedding_features": pluker, "sample_locs_dict": masks,
            "cond_frame_index": torch.zeros(1, dtype=torch.long), "add_type": "add_to_main_branch"}, F


def gen_unet_small(check):
    model = build_ref(SMALL_UNET, 128)
    unet = model.model.diffusion_model
    inp = synth_inputs(SMALL_CFG, SMALL_HW, 2, "small")
    cam, F = camera_condition(model, 8 * SMALL_HW, "pan_yaw", inp["pluker"])
    out = {"F": F.numpy()}
    t = torch.full((1,), 599, dtype=torch.long)
    xc = torch.cat([inp["x"], inp["c_concat"]], dim=1)
    with torch.no_grad():
        y_c = unet(xc, t, context=inp["ctx_cond"], fs=inp["fs"], camera_condition=cam)
        y_u = unet(xc, t, context=inp["ctx_uncond"], fs=inp["fs"], camera_condition=cam)
        y_n = unet(xc, t, context=inp["ctx_cond"], fs=inp["fs"], camera_condition=None)
    out.update(y_cond=y_c.numpy(), y_uncond=y_u.numpy(), y_nocam=y_n.numpy())
    print(f"  small unet: out std {y_c.std():.4f} / {y_u.std():.4f} / {y_n.std():.4f}")

    # one full CFG step through the reference's own sampler (ddim.py:241-346)
    DDIM = rh.patch_ddim_for_cpu()
    sampler = DDIM(model)
    sampler.make_schedule(25, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    for k in ("ddim_timesteps", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
        out["sched." + k] = np.asarray(getattr(sampler, k), dtype=np.float64 if k == "ddim_timesteps" else np.float32)
    out["sched.alphas_cumprod"] = model.alphas_cumprod.numpy()
    cond = {"c_crossattn": [inp["ctx_cond"]], "c_concat": [inp["c_concat"]], "camera_condition": cam}
    uc = {"c_crossattn": [inp["ctx_uncond"]], "c_concat": [inp["c_concat"]]}
    index = 14
    step = int(sampler.ddim_timesteps[index])
    ts = torch.full((1,), step, dtype=torch.long)
    torch.manual_seed(20230211)
    x_prev, pred_x0 = sampler.p_sample_ddim(inp["x"], cond, ts, index=index, unconditional_guidance_scale=3.5,
                                            unconditional_conditioning=uc, guidance_rescale=0.7, fs=inp["fs"],
                                            enable_camera_condition=True)
    out.update(step_x_prev=x_prev.numpy(), step_pred_x0=pred_x0.numpy(), step_index=np.int64(index), step_t=np.int64(step))
    np.savez_compressed(os.path.join(GOLD, "unet_small.npz"), **out)

    # the whole 25-step sampling loop of the reference (ddim.py:59-238): x_T given, eta-noise from torch's CPU generator
    # seeded right before the call (one randn per step, ddim.py:340), CFG 3.5, guidance_rescale 0.7, uniform_trailing
    import time
    t0 = time.time()
    torch.manual_seed(20230211)
    pred = []
    samples, _ = sampler.sample(25, 1, tuple(inp["x"].shape[1:]), conditioning=cond, eta=1.0, verbose=False, x_T=inp["x"],
                                unconditional_guidance_scale=3.5, unconditional_conditioning=uc, fs=inp["fs"],
                                timestep_spacing="uniform_trailing", guidance_rescale=0.7, enable_camera_condition=True,
                                img_callback=lambda p0, i: pred.append(p0.clone()))
    print(f"  25-step reference loop on the small model: {time.time() - t0:.1f} s, final std {samples.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "loop_small.npz"), x_final=samples.numpy(), pred_x0_step5=pred[4].numpy(),
                        pred_x0_step15=pred[14].numpy(), pred_x0_final=pred[-1].numpy(), seed=np.int64(20230211))

    if check:
        from oracle.unet_oracle import UNetOracle
        orc = UNetOracle(unet.state_dict(), SMALL_CFG)
        for name, ctx, c, ref in (("cond", inp["ctx_cond"], cam, y_c), ("uncond", inp["ctx_uncond"], cam, y_u), ("nocam", inp["ctx_cond"], None, y_n)):
            y = orc.forward(xc, t, ctx, inp["fs"], c)
            print(f"    oracle vs reference [{name}]: rel-L2 {rel_err(y, ref)[0]:.3e}  max|err|/max|ref| {rel_err(y, ref)[1]:.3e}")


def gen_unet_full(check):
    cfg = UNetConfig()
    t0 = time.time()
    model = build_ref(None, None)
    unet = model.model.diffusion_model
    print(f"  full model built+filled in {time.time() - t0:.0f}s")
    inp = synth_inputs(cfg, 32, 2, "full")
    cam, F = camera_condition(model, 256, "pan_yaw", inp["pluker"])
    t = torch.full((1,), 599, dtype=torch.long)
    xc = torch.cat([inp["x"], inp["c_concat"]], dim=1)
    out = {"F": F.numpy()}
    with torch.no_grad():
        t0 = time.time()
        y_c = unet(xc, t, context=inp["ctx_cond"], fs=inp["fs"], camera_condition=cam)
        print(f"  reference cond pass {time.time() - t0:.1f}s  out std {y_c.std():.4f}")
        t0 = time.time()
        y_u = unet(xc, t, context=inp["ctx_uncond"], fs=inp["fs"], camera_condition=cam)
        print(f"  reference uncond pass {time.time() - t0:.1f}s")
    out.update(y_cond=y_c.numpy(), y_uncond=y_u.numpy())
    np.savez_compressed(os.path.join(GOLD, "unet_full.npz"), **out)
    if check:
        from oracle.unet_oracle import UNetOracle
        orc = UNetOracle(unet.state_dict(), cfg)
        t0 = time.time()
        y = orc.forward(xc, t, inp["ctx_cond"], inp["fs"], cam)
        print(f"    oracle cond pass {time.time() - t0:.1f}s; vs reference rel-L2 {rel_err(y, y_c)[0]:.3e} max-norm {rel_err(y, y_c)[1]:.3e}")


def gen_loop_full(check, steps_saved=(1, 2, 5, 10, 15, 20, 25)):
    """The north star's acceptance case at FULL size: the reference's own DDIMSampler.sample (25 steps, uniform_trailing, eta 1,
    CFG 3.5, guidance_rescale 0.7, camera condition on both CFG branches) on the 1500.9 M-parameter UNet, 256x256x16f, batch 1,
    845-token context (1 reference + 2 context frames), epipolar masks at all four levels.  x_T = the `full` synthetic latent,
    eta-noise from torch's CPU generator seeded 20230211 right before the call (one randn per step, ddim.py:340).
    About 1 min of CPU per step on 8 cores (25 min in total).  The first saved step doubles as the golden of ONE full-size
    p_sample_ddim call (index 24, t = 999)."""
    cfg = UNetConfig()
    t0 = time.time()
    model = build_ref(None, None)
    print(f"  full model built+filled in {time.time() - t0:.0f}s")
    inp = synth_inputs(cfg, 32, 2, "full")
    cam, F = camera_condition(model, 256, "pan_yaw", inp["pluker"])
    DDIM = rh.patch_ddim_for_cpu()
    sampler = DDIM(model)
    cond = {"c_crossattn": [inp["ctx_cond"]], "c_concat": [inp["c_concat"]], "camera_condition": cam}
    uc = {"c_crossattn": [inp["ctx_uncond"]], "c_concat": [inp["c_concat"]]}
    out = {"F": F.numpy(), "seed": np.int64(20230211), "steps_saved": np.asarray(steps_saved, dtype=np.int64)}
    tt = [time.time()]

    def cb(i):
        tt.append(time.time())
        print(f"    step {i + 1}/25: {tt[-1] - tt[-2]:.1f}s", flush=True)

    torch.manual_seed(20230211)
    samples, inter = sampler.sample(25, 1, tuple(inp["x"].shape[1:]), conditioning=cond, eta=1.0, verbose=False, x_T=inp["x"],
                                    unconditional_guidance_scale=3.5, unconditional_conditioning=uc, fs=inp["fs"],
                                    timestep_spacing="uniform_trailing", guidance_rescale=0.7, enable_camera_condition=True,
                                    log_every_t=1, callback=cb)
    # intermediates[k] = state after k steps (entry 0 is x_T), ddim.py:156, 229-231
    assert len(inter["x_inter"]) == 26 and torch.equal(inter["x_inter"][-1], samples)
    for k in steps_saved:
        out[f"x_step{k}"] = inter["x_inter"][k].numpy()
        out[f"pred_x0_step{k}"] = inter["pred_x0"][k].numpy()
    out["sec_per_step"] = np.float64((tt[-1] - tt[0]) / 25)
    out["threads"] = np.int64(torch.get_num_threads())
    print(f"  25-step FULL-SIZE reference loop: {tt[-1] - tt[0]:.0f}s, final std {samples.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "loop_full.npz"), **out)


def gen_variants(check):
    """CameraCtrl / MotionCtrl baselines (R/baseline/*): one small-config UNet pass each through the reference's classes."""
    out = {}
    for kind in ("cameractrl", "motionctrl"):
        model = rh.build_baseline_model(kind, unet_overrides=SMALL_UNET)
        unet = model.model.diffusion_model
        synth.fill_module_(unet, seed=3)
        cfg = UNetConfig(model_channels=64, origin_h=128, origin_w=128, variant=kind)
        inp = synth_inputs(cfg, SMALL_HW, 0, "variant")
        xc = torch.cat([inp["x"], inp["c_concat"]], dim=1)
        t = torch.full((1,), 399, dtype=torch.long)
        if kind == "cameractrl":
            cam = {"pluker_embedding_features": inp["pluker"]}
        else:
            cam = {"RT": synth.synth_tensor("variant.RT", (1, 16, 12), 5)}
        with torch.no_grad():
            y = unet(xc, t, context=inp["ctx_uncond"], fs=inp["fs"], camera_condition=cam)
        out[f"{kind}.y"] = y.numpy()
        out[f"{kind}.nkeys"] = np.int64(len(unet.state_dict()))
        print(f"  {kind}: out std {y.std():.4f}, {len(unet.state_dict())} tensors")
        if check:
            from oracle.unet_oracle import UNetOracle
            yo = UNetOracle(unet.state_dict(), cfg).forward(xc, t, inp["ctx_uncond"], inp["fs"], cam)
            print(f"    oracle vs reference [{kind}]: rel-L2 {rel_err(yo, y)[0]:.3e} max-norm {rel_err(yo, y)[1]:.3e}")
            from camc2v_b200.modules import build_unet
            with torch.device("meta"):
                mine = {k: tuple(v.shape) for k, v in build_unet(cfg, variant=kind).state_dict().items()}
            ref = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
            print(f"    state_dict keys/shapes equal to camc2v_b200.modules: {mine == ref}")
    np.savez_compressed(os.path.join(GOLD, "variants.npz"), **out)


def gen_adaptor(check):
    """MultiLatentEpipolarAdaptor (SURVEY f-1) + conditional epipolar mask, reduced size: 8x8 latents, 2 context frames."""
    import json
    rh.setup_reference_imports()
    from model.modules.adaptors import MultiLatentEpipolarAdaptor
    kw = dict(query_dim=128, num_queries=64, video_length=16, embedding_dim=4, output_dim=4, depth=2, checkpoint=False,
              timestep_embedding_type="sinusoidal_embedded", use_plucker_embedding=False)
    torch.manual_seed(0)
    ref = MultiLatentEpipolarAdaptor(**kw).eval()
    synth.fill_module_(ref, seed=5)
    model = build_ref(SMALL_UNET, 128)                       # only for its (bound) camera-geometry methods
    hw_img, h = 64, 8
    K, w2c = synth.synth_camera("pan_yaw", T=16, H=hw_img, W=hw_img, B=1)
    _, w2c_o = synth.synth_camera("orbit", T=16, H=hw_img, W=hw_img, B=1)
    w2c_cond = w2c_o[:, [5, 11]].contiguous()                # two additional context views
    cond_idx = torch.zeros(1, dtype=torch.long)
    # compute_conditional_epipolar_mask (camcontexti2v.py:493-521) with the batch look-ups replaced by the tensors themselves
    from einops import rearrange, repeat
    c2w, c2w_cond = w2c.float().inverse(), w2c_cond.float().inverse()
    c2w_cond = torch.cat((c2w[torch.arange(1), cond_idx].unsqueeze(1), c2w_cond), dim=1)
    rel = model.get_pairwise_relative_pose(c2w_cond, c2w)
    rel = rearrange(rel, "B T C H W -> B C T H W")
    R, t = rel[..., :3, :3], rel[..., :3, 3:4]
    C = R.shape[2]
    Kr = repeat(K.float(), "B T H W -> B (T C) H W", C=C)
    Fm = model.get_fundamental_matrix(Kr, rearrange(R, "B T C H W -> B (T C) H W"), rearrange(t, "B T C H W -> B (T C) H W"))
    Fm = rearrange(Fm, "B (T C) H W -> B T C H W", C=C)
    mask = model.get_epipolar_mask(Fm, 16, h, h, 8, True)
    z = synth.synth_tensor("adaptor.z", (1, C * h * h, 4), 9)
    with torch.no_grad():
        y = ref(z, mask)
        y_nomask = ref(z, None)
    print(f"  adaptor: mask {tuple(mask.shape)} density {mask.float().mean():.4f}, out std {y.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "adaptor_small.npz"), K=K.numpy(), w2c=w2c.numpy(), w2c_cond=w2c_cond.numpy(), F=Fm.numpy(),
                        mask_packed=np.packbits(mask.numpy(), axis=-1), y=y.numpy(), y_nomask=y_nomask.numpy(),
                        kwargs=json.dumps(kw))
    json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, open(os.path.join(GOLD, "state_dict_adaptor.json"), "w"), indent=0)
    if check:
        import oracle
        from oracle import adaptor_oracle
        sd = ref.state_dict()
        Fo = adaptor_oracle.conditional_fundamental_matrices(K, w2c, w2c_cond, cond_idx)
        mo = oracle.epipolar_mask(Fo, h, h, 8)
        yo = adaptor_oracle.adaptor_forward(sd, z, mo, depth=2)
        print(f"    oracle F bit-identical: {bool(torch.equal(Fo, Fm))}; mask identical: {bool(torch.equal(mo, mask))}; "
              f"y rel-L2 {rel_err(yo, y)[0]:.3e}")


def gen_resampler(check):
    """Resampler image-token projector (SURVEY f-4), reduced size."""
    import json
    rh.setup_reference_imports()
    from lvdm.modules.encoders.resampler import Resampler
    kw = dict(dim=128, depth=2, dim_head=64, heads=4, num_queries=4, embedding_dim=96, output_dim=128, ff_mult=4, video_length=16,
              use_timestep_emb=True)
    torch.manual_seed(0)
    ref = Resampler(**kw).eval()
    synth.fill_module_(ref, seed=6)
    x = synth.synth_tensor("resampler.x", (2, 33, 96), 10)
    with torch.no_grad():
        y = ref(x)
    print(f"  resampler: out {tuple(y.shape)} std {y.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "resampler_small.npz"), y=y.numpy(), kwargs=json.dumps(kw))
    json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, open(os.path.join(GOLD, "state_dict_resampler.json"), "w"), indent=0)
    if check:
        from oracle import resampler_oracle
        yo = resampler_oracle.resampler_forward(ref.state_dict(), x, depth=2, heads=4)
        print(f"    oracle vs reference: rel-L2 {rel_err(yo, y)[0]:.3e}")


def gen_vae(check):
    """Decoder half of the first-stage AutoencoderKL (SURVEY f-3), reduced width: ch 64, two 8x8 latent frames -> 64x64 images."""
    import json
    rh.setup_reference_imports()
    from lvdm.modules.networks.ae_modules import Decoder
    dd = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
              attn_resolutions=[], dropout=0.0)
    torch.manual_seed(0)

    class Ref(torch.nn.Module):                               # AutoencoderKL.decode (autoencoder.py:103-106) without the encoder / loss
        def __init__(self):
            super().__init__()
            self.decoder = Decoder(**dd)
            self.post_quant_conv = torch.nn.Conv2d(4, dd["z_channels"], 1)

        def forward(self, z):
            return self.decoder(self.post_quant_conv(z))

    ref = Ref().eval()
    synth.fill_module_(ref, seed=7)
    z = synth.synth_tensor("vae.z", (2, 4, 8, 8), 11)
    with torch.no_grad():
        y = ref(z)
    print(f"  vae decoder: out {tuple(y.shape)} std {y.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "vae_small.npz"), y=y.numpy(), ddconfig=json.dumps(dd))
    json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, open(os.path.join(GOLD, "state_dict_vae_decoder.json"), "w"), indent=0)
    if check:
        from oracle import vae_oracle
        yo = vae_oracle.decode(ref.state_dict(), z, dd["ch_mult"], dd["num_res_blocks"])
        print(f"    oracle vs reference: rel-L2 {rel_err(yo, y)[0]:.3e}")

    # encoder half: AutoencoderKL.encode (autoencoder.py:97-101) up to the posterior moments
    from lvdm.modules.networks.ae_modules import Encoder

    class RefE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = Encoder(**dd)
            self.quant_conv = torch.nn.Conv2d(2 * dd["z_channels"], 2 * 4, 1)

        def forward(self, x):
            return self.quant_conv(self.encoder(x))

    torch.manual_seed(0)
    refe = RefE().eval()
    synth.fill_module_(refe, seed=8)
    x = synth.synth_tensor("vae.x", (2, 3, 64, 64), 12)
    with torch.no_grad():
        mom = refe(x)
    print(f"  vae encoder: moments {tuple(mom.shape)} std {mom.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "vae_enc_small.npz"), moments=mom.numpy(), ddconfig=json.dumps(dd))
    json.dump({k: list(v.shape) for k, v in refe.state_dict().items()}, open(os.path.join(GOLD, "state_dict_vae_encoder.json"), "w"), indent=0)
    if check:
        mo = vae_oracle.encode_moments(refe.state_dict(), x, dd["ch_mult"], dd["num_res_blocks"])
        print(f"    encoder oracle vs reference: rel-L2 {rel_err(mo, mom)[0]:.3e}")


POSE_KW = dict(downscale_factor=8, channels=[320, 640], nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False, compression_factor=1,
               temporal_attention_nhead=8, attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
               temporal_position_encoding_max_len=16)


# the class's other code paths: 3x3 in_conv / block2 convolutions (ksize 3), three blocks per level with the compression_factor
# bottleneck, two attention blocks per transformer, no position encoding.  (sk=False cannot run in the reference either unless
# every in_c == out_c: its skep conv is applied to the in_conv OUTPUT, camera_pose_encoder.py:257-266.)
POSE_KW_GENERIC = dict(downscale_factor=8, channels=[128, 256], nums_rb=3, cin=384, ksize=3, sk=True, use_conv=False, compression_factor=2,
                       temporal_attention_nhead=8, attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=False,
                       temporal_position_encoding_max_len=16)


def gen_pose_encoder(check):
    """CameraPoseEncoder (SURVEY f-2) in the shipped configuration (camcontexti2v_256.yaml:125-139) cut to its first two levels
    (head dims 40 and 80), on the Pluecker embedding of a 64 x 64 orbit trajectory.  The reference class itself runs; the two
    `diffusers` classes it imports are the restated stand-ins of oracle/refgen/diffusers_stub.py (diffusers is not installed)."""
    import json
    import oracle
    from oracle import camera_oracle, pose_encoder_oracle
    import diffusers_stub
    diffusers_stub.install()
    rh.setup_reference_imports()
    from model.modules.camera_pose_encoder import CameraPoseEncoder
    torch.manual_seed(0)
    ref = CameraPoseEncoder(**POSE_KW).eval()
    pe = {k: v.clone() for k, v in ref.state_dict().items() if k.endswith("pos_encoder.pe")}
    synth.fill_module_(ref, seed=8)
    ref.load_state_dict(pe, strict=False)                    # keep the sinusoidal buffers the constructor made
    K, w2c = synth.synth_camera("orbit", T=16, H=64, W=64, B=1)
    rel = camera_oracle.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))
    x = oracle.plucker(K, rel, 64, 64, "plucker")            # == CameraControlLVDM.ray_condition (tests/test_oracle_golden.py)
    with torch.no_grad():
        feats = ref(x)
    print("  pose encoder:", [tuple(f.shape) for f in feats], [round(float(f.std()), 4) for f in feats])
    np.savez_compressed(os.path.join(GOLD, "pose_encoder_small.npz"), **{f"f{i}": f.numpy() for i, f in enumerate(feats)},
                        kwargs=json.dumps(POSE_KW))
    json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, open(os.path.join(GOLD, "state_dict_pose_encoder.json"), "w"), indent=0)
    for k, v in pe.items():
        d = v.shape[-1]
        assert torch.equal(v[0], pose_encoder_oracle.positional_encoding(d, 16)), k
    if check:
        fo = pose_encoder_oracle.pose_encoder_forward(ref.state_dict(), x, n_levels=2)
        print("    oracle vs reference rel-L2:", [f"{rel_err(a, b)[0]:.3e}" for a, b in zip(fo, feats)])
    torch.manual_seed(0)
    ref2 = CameraPoseEncoder(**POSE_KW_GENERIC).eval()
    synth.fill_module_(ref2, seed=9)
    with torch.no_grad():
        feats2 = ref2(x)
    print("  pose encoder (generic options):", [tuple(f.shape) for f in feats2], [round(float(f.std()), 4) for f in feats2])
    np.savez_compressed(os.path.join(GOLD, "pose_encoder_generic.npz"), **{f"f{i}": f.numpy() for i, f in enumerate(feats2)},
                        kwargs=json.dumps(POSE_KW_GENERIC))
    json.dump({k: list(v.shape) for k, v in ref2.state_dict().items()}, open(os.path.join(GOLD, "state_dict_pose_encoder_generic.json"), "w"),
              indent=0)
    if check:
        fo = pose_encoder_oracle.pose_encoder_forward(ref2.state_dict(), x, n_levels=2, nums_rb=3, n_attn=2)
        print("    oracle vs reference rel-L2:", [f"{rel_err(a, b)[0]:.3e}" for a, b in zip(fo, feats2)])


def gen_camcfg(check):
    """p_sample_ddim with camera guidance (camera_cfg = 2, cosine scheduler: a third UNet pass, ddim.py:268-280) on the small model."""
    model = build_ref(SMALL_UNET, 128)
    inp = synth_inputs(SMALL_CFG, SMALL_HW, 2, "small")
    cam, F = camera_condition(model, 8 * SMALL_HW, "pan_yaw", inp["pluker"])
    DDIM = rh.patch_ddim_for_cpu()
    sampler = DDIM(model)
    sampler.make_schedule(25, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    cond = {"c_crossattn": [inp["ctx_cond"]], "c_concat": [inp["c_concat"]], "camera_condition": cam}
    uc = {"c_crossattn": [inp["ctx_uncond"]], "c_concat": [inp["c_concat"]]}
    index = 9
    step = int(sampler.ddim_timesteps[index])
    ts = torch.full((1,), step, dtype=torch.long)
    torch.manual_seed(20230211)
    x_prev, pred_x0 = sampler.p_sample_ddim(inp["x"], cond, ts, index=index, unconditional_guidance_scale=3.5, unconditional_conditioning=uc,
                                            guidance_rescale=0.7, fs=inp["fs"], enable_camera_condition=True, camera_cfg=2.0,
                                            camera_cfg_scheduler="cosine")
    print(f"  camera_cfg step: x_prev std {x_prev.std():.4f}")
    np.savez_compressed(os.path.join(GOLD, "camcfg_small.npz"), x_prev=x_prev.numpy(), pred_x0=pred_x0.numpy(), index=np.int64(index),
                        t=np.int64(step), camera_cfg=np.float64(2.0))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--check-oracle", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if a.only in (None, "masks"):
        print("[masks]")
        gen_masks(rh.build_reference_model(unet_overrides=SMALL_UNET), a.check_oracle)
    if a.only in (None, "unet_small", "ddim"):
        print("[unet_small + ddim step]")
        gen_unet_small(a.check_oracle)
    if a.only in (None, "variants"):
        print("[variants]")
        gen_variants(a.check_oracle)
    if a.only in (None, "adaptor"):
        print("[adaptor]")
        gen_adaptor(a.check_oracle)
    if a.only in (None, "resampler"):
        print("[resampler]")
        gen_resampler(a.check_oracle)
    if a.only in (None, "vae"):
        print("[vae]")
        gen_vae(a.check_oracle)
    if a.only in (None, "pose_encoder"):
        print("[pose_encoder]")
        gen_pose_encoder(a.check_oracle)
    if a.only in (None, "camcfg"):
        print("[camcfg]")
        gen_camcfg(a.check_oracle)
    if a.only in ("unet_full",):
        print("[unet_full]")
        gen_unet_full(a.check_oracle)
    if a.only in ("loop_full",):
        print("[loop_full]")
        gen_loop_full(a.check_oracle)
