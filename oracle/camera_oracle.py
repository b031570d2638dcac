"""CPU ORACLE (test infrastructure, NOT product code): camera geometry feeding the epipolar mask.

Restates (torch fp32 on CPU, same op order as the reference so that F is bit-identical on the same host):
  * get_batch_input_camera_condition_process   R/model/camcontexti2v.py:525-554
  * get_relative_pose (mode='left')            R/model/base.py:176-198
  * get_relative_c2w_RT_pairs                  R/model/camcontexti2v.py:174-184
  * add_small_perturbation                     R/model/camcontexti2v.py:273-278   (consumes torch RNG!)
  * get_fundamental_matrix                     R/model/camcontexti2v.py:188-198
"""
from __future__ import annotations

import torch

from . import epipolar_mask


def relative_c2w(w2c: torch.Tensor, cond_frame_index: torch.Tensor, trace_scale_factor: float = 1.0) -> torch.Tensor:
    c2w = w2c.float().inverse()
    b = c2w.shape[0]
    first = c2w[torch.arange(b), cond_frame_index].unsqueeze(1)
    rel = first.inverse() @ c2w
    rel[:, :, :3, 3] = rel[:, :, :3, 3] * trace_scale_factor
    return rel


def fundamental_matrices(K: torch.Tensor, rel_c2w: torch.Tensor, perturb: bool = True, eps: float = 1e-6) -> torch.Tensor:
    """K [B,T,3,3], rel_c2w [B,T,4,4] -> F [B,T1,T2,3,3].  Draws randn from the global torch RNG iff perturb."""
    inv = rel_c2w.inverse()[:, None]                 # b 1 t
    pairs = inv @ rel_c2w[:, :, None]                # b t1 t2 : inv(rel[t2]) @ rel[t1]
    R = pairs[..., :3, :3]
    t = pairs[..., :3, 3:4]
    if perturb:
        zero = (t.abs() < eps).all(dim=-2, keepdim=True)
        noise = torch.randn_like(t) * eps
        t = torch.where(zero, noise, t)
    Kb = K.float().unsqueeze(1)
    E = torch.cross(t, R, dim=-2)
    Kinv = torch.inverse(Kb)
    return Kinv.transpose(-1, -2) @ E @ Kinv


def camera_condition_masks(K, w2c, cond_frame_index, H=256, W=256, resolutions=(8, 4, 2, 1), trace_scale_factor=1.0, perturb=True):
    """Returns (F, {d: bool mask [B, T*h*w, T*h*w]}, rel_c2w) with d = 8*ds for ds in resolutions."""
    rel = relative_c2w(w2c, cond_frame_index, trace_scale_factor)
    Fm = fundamental_matrices(K, rel, perturb)
    masks = {int(8 * ds): epipolar_mask(Fm, H // int(8 * ds), W // int(8 * ds), int(8 * ds)) for ds in resolutions}
    return Fm, masks, rel
