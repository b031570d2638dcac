"""CPU timing of the UNMODIFIED reference on the bench workload: whole `DDIMSampler.p_sample_ddim` CFG steps.

MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `--impl reference` arm and `cpu_baseline` leg).  The code that runs is the
reference's own (R/lvdm/models/samplers/ddim.py:241-346 -> LatentDiffusion.apply_model ddpm3d.py:724-739 ->
new_forward_for_unet modified_forwards.py:29-131 and every module below it), imported by ref_harness.py from
/root/reference (build container) or from the byte-identical copies under oracle/_ref/ (GPU box; install_ref.py).
Nothing of camc2v_b200's kernels, modules or sampler is on that path; from this repo it only takes the seeded synthetic
weights / inputs (camc2v_b200.synth, .testing) so that both arms of the bench see the same workload:
CamContextI2V UNet 1500.9 M params, B=1, 4x16x32x32 latent, 845-token cond / 333-token uncond context, epipolar masks at
all four levels (built by the reference's own get_epipolar_mask), CFG 3.5, guidance_rescale 0.7, eta 1, uniform_trailing.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def reference_available() -> str | None:
    """None when the reference can be imported here, else the reason (one line)."""
    import ref_harness as rh
    if not os.path.isdir(rh.REF_PKG):
        return f"reference sources not found ({rh.REF_PKG}); run oracle/refgen/install_ref.py in the build container"
    return None


class ReferenceStepper:
    """Consecutive DDIM steps of the reference's sampler on the bench workload (x <- x_prev after every step)."""

    def __init__(self, threads: int | None = None, tag: str = "bench0"):
        import ref_harness as rh
        from camc2v_b200 import synth
        from camc2v_b200.config import UNetConfig
        from camc2v_b200.testing import synth_unet_inputs

        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        t0 = time.perf_counter()
        cfg = UNetConfig()
        self.model = rh.build_reference_model()
        synth.fill_module_(self.model.model.diffusion_model, seed=0)
        inp = synth_unet_inputs(cfg, 32, 2, tag, B=1)
        K, w2c = synth.synth_camera("pan_yaw", T=cfg.temporal_length, B=1)
        torch.manual_seed(123)
        m = self.model
        with torch.no_grad():                  # geometry half of get_batch_input_camera_condition_process (camcontexti2v.py:525-554)
            c2w = w2c.float().inverse()
            rel = m.get_relative_pose(c2w, torch.zeros(1, dtype=torch.long), mode="left", normalize_T0=False)
            pairs = m.get_relative_c2w_RT_pairs(rel)
            R, t = pairs[..., :3, :3], pairs[..., :3, 3:4]
            t = m.add_small_perturbation(t, epsilon=1e-6)
            F = m.get_fundamental_matrix(K.float().unsqueeze(1), R, t)
            masks = {int(8 * ds): m.get_epipolar_mask(F, 16, 256 // int(8 * ds), 256 // int(8 * ds), int(8 * ds)) for ds in (8, 4, 2, 1)}
        cam = {"pluker_embedding_features": inp["pluker"], "sample_locs_dict": masks,
               "cond_frame_index": torch.zeros(1, dtype=torch.long), "add_type": "add_to_main_branch"}
        DDIM = rh.patch_ddim_for_cpu()
        self.sampler = DDIM(self.model)
        self.sampler.make_schedule(25, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
        self.cond = {"c_crossattn": [inp["ctx_cond"]], "c_concat": [inp["c_concat"]], "camera_condition": cam}
        self.uc = {"c_crossattn": [inp["ctx_uncond"]], "c_concat": [inp["c_concat"]]}
        self.fs = inp["fs"]
        self.x0 = inp["x"]
        self.x = inp["x"]
        self.i = 0
        self.ts = np.flip(self.sampler.ddim_timesteps).copy()
        self.setup_s = time.perf_counter() - t0
        torch.manual_seed(20230211)

    def step(self) -> float:
        """One `p_sample_ddim` (cond pass + uncond pass + CFG combine + rescale + DDIM update); returns its wall time in s."""
        i = self.i % 25
        if i == 0:
            self.x = self.x0
        index = 24 - i
        ts = torch.full((1,), int(self.ts[i]), dtype=torch.long)
        t0 = time.perf_counter()
        with torch.no_grad():
            x_prev, _ = self.sampler.p_sample_ddim(self.x, self.cond, ts, index=index, unconditional_guidance_scale=3.5,
                                                   unconditional_conditioning=self.uc, guidance_rescale=0.7, fs=self.fs,
                                                   enable_camera_condition=True)
        dt = time.perf_counter() - t0
        self.x = x_prev
        self.i += 1
        return dt


SAMPLE_TEXT = ("whole DDIMSampler.p_sample_ddim CFG steps of the UNMODIFIED reference (cond + uncond UNet pass at full size, B=1, 4x16x32x32 "
               "latent, 845 / 333-token contexts, epipolar masks at 4 levels, CFG 3.5, rescale 0.7, eta 1; fp32, torch CPU, einsum / SDPA "
               "attention fallbacks the reference itself selects without xformers) - no extrapolation")
