"""CPU ORACLE (test infrastructure, NOT product code): DDIM schedule + one CFG p_sample_ddim update.

Restates
  * make_beta_schedule('linear')                 R/lvdm/models/utils_diffusion.py:31-36
  * DDPM.register_schedule alphas_cumprod        R/lvdm/models/ddpm3d.py:125-188
  * make_ddim_timesteps / sampling parameters    R/lvdm/models/utils_diffusion.py:56-91
  * DDIMSampler.p_sample_ddim (CFG, rescale, update)  R/lvdm/models/samplers/ddim.py:241-346
  * rescale_noise_cfg                            R/lvdm/models/utils_diffusion.py:147-158
Known-answer values (25 steps, uniform_trailing, eta=1): t = [39, 79, ..., 999],
ddim_alphas[0..2] = 0.96289510, 0.91947132, 0.86999440, ddim_alphas[-1] = 0.00466010,
ddim_sigmas[0] = 0.02883150, ddim_sigmas[-1] = 0.61106440.
"""
from __future__ import annotations

import numpy as np
import torch


def alphas_cumprod(timesteps=1000, linear_start=0.00085, linear_end=0.012) -> np.ndarray:
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas, axis=0).astype(np.float32)


def ddim_timesteps(method: str, n_ddim: int, n_ddpm: int = 1000) -> np.ndarray:
    if method == "uniform":
        c = n_ddpm // n_ddim
        return np.asarray(list(range(0, n_ddpm, c))) + 1
    if method == "uniform_trailing":
        c = n_ddpm / n_ddim
        return np.flip(np.round(np.arange(n_ddpm, 0, -c))).astype(np.int64) - 1
    raise NotImplementedError(method)


def ddim_schedule(n_ddim=25, eta=1.0, method="uniform_trailing", n_ddpm=1000):
    """Returns dict of float32 arrays: timesteps, alphas, alphas_prev, sigmas, sqrt_one_minus_alphas."""
    ac = alphas_cumprod(n_ddpm)
    ts = ddim_timesteps(method, n_ddim, n_ddpm)
    a = ac[ts].astype(np.float64)
    a_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist(), dtype=np.float64)
    sig = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return dict(timesteps=ts, alphas=a.astype(np.float32), alphas_prev=a_prev.astype(np.float32),
                sigmas=sig.astype(np.float32), sqrt_one_minus_alphas=np.sqrt(1.0 - a).astype(np.float32))


def cfg_ddim_update(x, e_cond, e_uncond, noise, a_t, a_prev, sigma_t, sqrt_one_minus_at, scale, guidance_rescale, e_cond_nocam=None,
                    cam_weight=0.0):
    """All tensors [B,C,T,H,W] fp32; scalars are python floats (fp32 values).  Returns (x_prev, pred_x0).
    e_cond_nocam / cam_weight: the camera guidance term of ddim.py:268-280, cam_weight = (camera_cfg - 1) * scheduler weight."""
    x, e_cond, e_uncond, noise = (t.float() for t in (x, e_cond, e_uncond, noise))
    e = e_uncond + scale * (e_cond - e_uncond)
    if e_cond_nocam is not None:
        e = e + cam_weight * (e_cond - e_cond_nocam.float())
    if guidance_rescale > 0.0:
        dims = list(range(1, e.ndim))
        std_c = e_cond.std(dim=dims, keepdim=True)
        std_e = e.std(dim=dims, keepdim=True)
        e = guidance_rescale * (e * (std_c / std_e)) + (1 - guidance_rescale) * e
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    pred_x0 = (x - f32(sqrt_one_minus_at) * e) / f32(a_t).sqrt()
    dir_xt = (1.0 - f32(a_prev) - f32(sigma_t) ** 2).clamp(min=0).sqrt() * e
    x_prev = f32(a_prev).sqrt() * pred_x0 + dir_xt + f32(sigma_t) * noise
    return x_prev, pred_x0
