"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

Every tensor is drawn from its own CPU generator seeded by (seed, crc32(name)), so the SAME values are
produced for the reference model (in the build container), the CPU oracle and the CUDA path (on the
GPU box) without shipping multi-GB fixtures.  Zero-initialised layers of the reference
(`proj_out`, `epipolar.to_out`, `pluker_projection`, ResBlock `out_layers[-1]`, `conv4`, `out[-1]`,
`fps_embedding[-1]`) are therefore re-randomised, otherwise the network is a no-op.
"""
from __future__ import annotations

import zlib

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2**62))
    return g


def synth_tensor(name: str, shape, seed: int = 0, std: float = 1.0) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(seed, name), dtype=torch.float32) * std


def synth_param(name: str, shape, seed: int = 0) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "alpha":  # CrossAttention.alpha (attention.py:83)
        return torch.randn(shape, generator=g) * 0.3
    if leaf == "register_tokens":  # epipolar.py:69
        return torch.randn(shape, generator=g)
    if leaf == "bias":
        return torch.randn(shape, generator=g) * 0.05
    if len(shape) == 1:  # GroupNorm / LayerNorm scale
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=g) * (fan_in ** -0.5)


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {key: shape}.  Returns {key: fp32 CPU tensor}."""
    return {k: synth_param(k, s, seed) for k, s in shapes.items()}


def fill_module_(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter/buffer of `module` with its synthetic value (in place)."""
    sd = module.state_dict()
    new = {k: synth_param(k, v.shape, seed).to(v.dtype) for k, v in sd.items()}
    module.load_state_dict(new, strict=True)


# ----------------------------------------------------------------------------------------------
# synthetic camera trajectories; w2c [T,4,4] + pixel-unit intrinsics [T,3,3] in the
# format of R/data/single_image_for_inference.py:111-117 (fx = 0.5*W, cx = 0.5*W).
# ----------------------------------------------------------------------------------------------
def _rot_y(a):
    c, s = torch.cos(a), torch.sin(a)
    z, o = torch.zeros_like(a), torch.ones_like(a)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def _rot_z(a):
    c, s = torch.cos(a), torch.sin(a)
    z, o = torch.zeros_like(a), torch.ones_like(a)
    return torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, o], -1)], -2)


def synth_camera(kind: str = "pan_yaw", T: int = 16, H: int = 256, W: int = 256, B: int = 1):
    """Returns (K [B,T,3,3], w2c [B,T,4,4]) fp32."""
    i = torch.arange(T, dtype=torch.float32)
    R = torch.eye(3).repeat(T, 1, 1)
    t = torch.zeros(T, 3)
    if kind == "pan_yaw":  # pan right + yaw + slight dolly (the survey's probe trajectory)
        R = _rot_y(0.02 * i)
        t = torch.stack([0.05 * i, torch.zeros(T), 0.01 * i], -1)
    elif kind == "stationary":  # all-zero translation: every pair gets the 1e-6 perturbation
        pass
    elif kind == "dolly":
        t = torch.stack([torch.zeros(T), torch.zeros(T), 0.06 * i], -1)
    elif kind == "yaw":
        R = _rot_y(0.03 * i)
    elif kind == "roll_pan_up":
        R = _rot_z(0.02 * i)
        t = torch.stack([torch.zeros(T), -0.04 * i, torch.zeros(T)], -1)
    elif kind == "orbit":
        R = _rot_y(-0.04 * i)
        t = torch.stack([torch.sin(0.04 * i) * 1.0, torch.zeros(T), 1.0 - torch.cos(0.04 * i)], -1)
    else:
        raise ValueError(kind)
    w2c = torch.eye(4).repeat(T, 1, 1)
    w2c[:, :3, :3] = R
    w2c[:, :3, 3] = t
    K = torch.tensor([[0.5 * W, 0.0, 0.5 * W], [0.0, 0.5 * W, 0.5 * H], [0.0, 0.0, 1.0]]).repeat(T, 1, 1)
    return K.unsqueeze(0).repeat(B, 1, 1, 1).contiguous(), w2c.unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
