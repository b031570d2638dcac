"""Drop-in proof: this package's UNet inside the UNMODIFIED reference, driven by the reference's own code.

The reference (byte-identical copies under oracle/_ref, installed by oracle/refgen/install_ref.py; /root/reference in the build
container) constructs its own `CamContextI2V` from its shipped YAML with ONE change - `unet_config.target:
camc2v_b200.modules.UNetModel`.  Its constructor then does what it always does
(R/model/camcontexti2v.py:111-170): re-binds forwards by class name and injects `Epipolar` sub-modules; the package declines
the re-binding and adopts the injected modules (camc2v_b200.modules._RefBindable, BasicTransformerBlock.add_module).  The
reference UNet's `state_dict` is loaded with strict=True, and the sample is produced

  (1) by the reference's own `DDIMSampler` -> `LatentDiffusion.apply_model` -> `DiffusionWrapper` (ddim.py:59-346,
      ddpm3d.py:724-739, 1268-1272) with the reference's own camera condition format (bool `sample_locs_dict` masks from its own
      `get_epipolar_mask`), and
  (2) by `camc2v_b200.sampler.DDIMSampler` on the same reference model object,

both compared with the golden of the pure reference (tests/golden/loop_full.npz).  The only harness-side touches: the stub
`pytorch_lightning`, Identity VAE / CLIP stages, and `noise_like` drawing the eta-noise from torch's CPU generator so that the
25 draws equal those of the CPU golden run.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"
TOL = float(os.environ.get("C2V_TEST_TOL", "5e-3"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refgen"))


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def dropin():
    import ref_harness as rh
    from camc2v_b200 import synth
    from camc2v_b200.config import UNetConfig
    from camc2v_b200.modules import UNetModel
    from camc2v_b200.testing import synth_unet_inputs
    if not os.path.isdir(rh.REF_PKG):
        pytest.skip("reference sources not installed (python oracle/refgen/install_ref.py in the build container)")
    torch.set_num_threads(os.cpu_count() or 1)
    # the pure reference UNet: the source of the state_dict
    ref = rh.build_reference_model()
    synth.fill_module_(ref.model.diffusion_model, seed=0)
    sd = ref.model.diffusion_model.state_dict()
    # the reference model with the UNet swapped by YAML target; everything else is the reference's constructor
    model = rh.build_reference_model(unet_target="camc2v_b200.modules.UNetModel")
    unet = model.model.diffusion_model
    assert isinstance(unet, UNetModel) and "new_forward_for_unet" in unet.__dict__.get("_declined_rebinds", [])
    missing = unet.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # camera condition in the reference's own format, from the reference's own geometry code (camcontexti2v.py:525-554)
    cfg = UNetConfig()
    inp = synth_unet_inputs(cfg, 32, 2, "full")
    K, w2c = synth.synth_camera("pan_yaw", T=16, H=256, W=256, B=1)
    torch.manual_seed(123)
    with torch.no_grad():
        c2w = w2c.float().inverse()
        relp = ref.get_relative_pose(c2w, torch.zeros(1, dtype=torch.long), mode="left", normalize_T0=False)
        pairs = ref.get_relative_c2w_RT_pairs(relp)
        t = ref.add_small_perturbation(pairs[..., :3, 3:4], epsilon=1e-6)
        F = ref.get_fundamental_matrix(K.float().unsqueeze(1), pairs[..., :3, :3], t)
        masks = {int(8 * ds): ref.get_epipolar_mask(F, 16, 256 // int(8 * ds), 256 // int(8 * ds), int(8 * ds)) for ds in (8, 4, 2, 1)}
    gl = np.load(os.path.join(GOLD, "loop_full.npz"))
    assert np.array_equal(gl["F"], F.numpy()), "F differs from the golden run's (CPU torch build mismatch?)"
    del ref, sd
    model = model.to(DEV).eval()
    cam = {"pluker_embedding_features": [p.to(DEV) for p in inp["pluker"]], "sample_locs_dict": {d: m.to(DEV) for d, m in masks.items()},
           "cond_frame_index": torch.zeros(1, dtype=torch.long, device=DEV), "add_type": "add_to_main_branch"}
    cond = {"c_crossattn": [inp["ctx_cond"].to(DEV)], "c_concat": [inp["c_concat"].to(DEV)], "camera_condition": cam}
    uc = {"c_crossattn": [inp["ctx_uncond"].to(DEV)], "c_concat": [inp["c_concat"].to(DEV)]}
    return model, cond, uc, inp, gl


def test_reference_sampler_drives_the_cuda_unet(dropin):
    """(1): the reference's DDIMSampler.sample, 25 steps, through the reference's apply_model into the swapped UNet."""
    model, cond, uc, inp, gl = dropin
    import lvdm.models.samplers.ddim as ref_ddim
    saved = ref_ddim.noise_like
    ref_ddim.noise_like = lambda shape, device, repeat=False: torch.randn(shape).to(device)      # CPU generator: the golden's draws
    try:
        sampler = ref_ddim.DDIMSampler(model)
        torch.manual_seed(int(gl["seed"]))
        samples, inter = sampler.sample(25, 1, tuple(inp["x"].shape[1:]), conditioning=cond, eta=1.0, verbose=False, x_T=inp["x"].to(DEV),
                                        unconditional_guidance_scale=3.5, unconditional_conditioning=uc, fs=inp["fs"].to(DEV),
                                        timestep_spacing="uniform_trailing", guidance_rescale=0.7, enable_camera_condition=True, log_every_t=1)
    finally:
        ref_ddim.noise_like = saved
    errs = {k: rel(inter["x_inter"][k], torch.from_numpy(gl[f"x_step{k}"])) for k in (1, 5, 15, 25)}
    print("reference DDIMSampler + CUDA UNet vs pure reference, x rel-L2 / max-norm: " +
          "; ".join(f"step {k}: {e[0]:.2e} / {e[1]:.2e}" for k, e in errs.items()))
    assert torch.isfinite(samples).all()
    for k, e in errs.items():
        assert e[0] < TOL and e[1] < TOL, (k, e)


def test_cuda_sampler_runs_on_the_reference_model(dropin):
    """(2): camc2v_b200.sampler.DDIMSampler on the reference's LatentDiffusion object (apply_model, alphas_cumprod, ...)."""
    from camc2v_b200.sampler import DDIMSampler
    model, cond, uc, inp, gl = dropin
    s = DDIMSampler(model)
    s.make_schedule(25, "uniform_trailing", 1.0, verbose=False)
    kw = dict(unconditional_guidance_scale=3.5, unconditional_conditioning=uc, guidance_rescale=0.7, fs=inp["fs"].to(DEV),
              enable_camera_condition=True)
    torch.manual_seed(int(gl["seed"]))
    x = inp["x"].to(DEV)
    errs = {}
    for i, step in enumerate(np.flip(s.ddim_timesteps)):
        ts = torch.full((1,), int(step), dtype=torch.long, device=DEV)
        noise = torch.randn(inp["x"].shape)
        x, p0 = s.p_sample_ddim(x, cond, ts, index=24 - i, noise=noise.to(DEV), **kw)
        if i + 1 in (1, 5, 15, 25):
            errs[i + 1] = rel(x, torch.from_numpy(gl[f"x_step{i + 1}"]))
    print("CUDA DDIMSampler on the reference model vs pure reference, x rel-L2 / max-norm: " +
          "; ".join(f"step {k}: {e[0]:.2e} / {e[1]:.2e}" for k, e in errs.items()))
    for k, e in errs.items():
        assert e[0] < TOL and e[1] < TOL, (k, e)
