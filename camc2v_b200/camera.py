"""Camera conditioning producers that feed the per-step path.

Mirrors the geometry half of `CamContextI2V.get_batch_input_camera_condition_process`
(R/model/camcontexti2v.py:525-572): w2c -> relative c2w poses -> pairwise relative poses -> (perturbed)
translations -> fundamental matrices -> epipolar masks, and the Pluecker / ray embedding
(R/model/base.py:112-174).

The 4x4 / 3x3 algebra on T=16 poses is host-side glue done with torch on whatever device the poses
live on, in the reference's op order (so that the torch RNG draw of `add_small_perturbation` and the bits
of F agree with the reference when run on the same device).  The heavy outputs are produced by the CUDA
library: `ops.plucker` ([B,6,T,256,256]) and — only if a caller insists on the reference's
`sample_locs_dict` format — `ops.epipolar_mask`.  The hot path itself never materialises a mask: it
consumes `epipolar_F` ([B,T,T,3,3], 9 floats per frame pair) inside the attention kernel.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops


def relative_c2w(w2c: torch.Tensor, cond_frame_index: torch.Tensor, trace_scale_factor: float = 1.0) -> torch.Tensor:
    """get_relative_pose(mode='left') of base.py:176-198 applied to c2w = inverse(w2c); camcontexti2v.py:533-538."""
    c2w = torch.linalg.inv(w2c.float())
    b = c2w.shape[0]
    ref = c2w[torch.arange(b, device=c2w.device), cond_frame_index.to(c2w.device)].unsqueeze(1)
    rel = torch.linalg.inv(ref) @ c2w
    rel[:, :, :3, 3] = rel[:, :, :3, 3] * trace_scale_factor
    return rel


def fundamental_matrices(K: torch.Tensor, rel_c2w: torch.Tensor, perturb: bool = True, eps: float = 1e-6) -> torch.Tensor:
    """camcontexti2v.py:174-198, 273-278, 540-549.  K [B,T,3,3], rel_c2w [B,T,4,4] -> F [B,T1,T2,3,3].
    With `perturb` the zero translations (t1 == t2, or a static camera) are replaced by randn*eps drawn from
    the global torch generator at this point, as `add_small_perturbation` does."""
    pairs = torch.linalg.inv(rel_c2w)[:, None] @ rel_c2w[:, :, None]
    R = pairs[..., :3, :3]
    t = pairs[..., :3, 3:4]
    if perturb:
        zero = (t.abs() < eps).all(dim=-2, keepdim=True)
        t = torch.where(zero, torch.randn_like(t) * eps, t)
    E = torch.linalg.cross(t.expand_as(R), R, dim=-2)
    Kinv = torch.linalg.inv(K.float().unsqueeze(1))
    return Kinv.transpose(-1, -2) @ E @ Kinv


def camera_condition(K: torch.Tensor, w2c: torch.Tensor, cond_frame_index: torch.Tensor, H: int = 256, W: int = 256,
                     pluker_embedding_features: Optional[Sequence[torch.Tensor]] = None, attention_resolution=(8, 4, 2, 1),
                     trace_scale_factor: float = 1.0, perturb: bool = True, materialize_masks: bool = False,
                     add_type: str = "add_to_main_branch", camera_embedding: str = "plucker", device="cuda") -> Dict:
    """Build the `camera_condition` dict consumed by UNetModel.forward.

    Same keys as the reference (`pluker_embedding_features`, `sample_locs_dict`, `cond_frame_index`, `add_type`)
    plus `epipolar_F` (the kernel-native form of the mask) and `pluker_embedding` (input of the pose encoder,
    which is once-per-sample and out of the per-step scope: SURVEY f-2)."""
    rel = relative_c2w(w2c, cond_frame_index, trace_scale_factor)
    Fm = fundamental_matrices(K, rel, perturb).to(device).contiguous()
    out = {
        "pluker_embedding_features": pluker_embedding_features,
        "epipolar_F": Fm,
        "sample_locs_dict": None,
        "cond_frame_index": cond_frame_index,
        "add_type": add_type,
        "pluker_embedding": ops.plucker(K.to(device), rel.to(device), H, W, camera_embedding),
        "relative_c2w": rel,
    }
    if materialize_masks:
        out["sample_locs_dict"] = {int(8 * ds): ops.epipolar_mask(Fm, H // int(8 * ds), W // int(8 * ds), int(8 * ds))
                                   for ds in attention_resolution}
    return out


def conditional_fundamental_matrices(K: torch.Tensor, w2c: torch.Tensor, w2c_cond: torch.Tensor, cond_frame_index: Optional[torch.Tensor]):
    """F [B, T, C, 3, 3] between the T target frames and the C = 1 + n context frames (reference frame first) for the adaptor's
    conditional epipolar mask (compute_conditional_epipolar_mask, R/model/camcontexti2v.py:493-521; get_pairwise_relative_pose,
    R/model/base.py:200-217).  Tiny 4x4 algebra, done with torch on the host like the rest of this file; the mask itself is
    ops.epipolar_mask(F, h, w, d)."""
    c2w = w2c.float().inverse()
    c2w_cond = w2c_cond.float().inverse()
    if cond_frame_index is not None:
        c2w_cond = torch.cat((c2w[torch.arange(len(cond_frame_index)), cond_frame_index].unsqueeze(1), c2w_cond), dim=1)
    rel = (c2w_cond.inverse()[:, :, None] @ c2w[:, None]).transpose(1, 2)          # [B, T, C]: inv(c2w_cond[c]) @ c2w[t]
    R, t = rel[..., :3, :3], rel[..., :3, 3:4]
    Kinv = torch.inverse(K.float()[:, :, None])
    return Kinv.transpose(-1, -2) @ torch.cross(t, R, dim=-2) @ Kinv
