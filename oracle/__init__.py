"""CPU ORACLE package — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Holds a CPU restatement of the reference's algorithm for the DDIM/CFG denoising hot path
(SURVEY.md §8c).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  `camc2v_b200/` (the product) never does, and fails loudly if
its CUDA library is missing rather than falling back to anything in here.

Parity pinning: the reference ships no tests / golden vectors.  The oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container by
`oracle/refgen/make_golden.py` (imports /root/reference unmodified) and committed under
`tests/golden/`; `tests/test_oracle_golden.py` re-checks the oracle against those files on any host.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libepi_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc-compile the C restatement (epi_oracle.c) next to its source."""
    src = os.path.join(_HERE, "epi_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.epi_mask_oracle.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
        _lib.epi_mask_oracle.restype = None
        _lib.epi_mask_oracle_rect.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_void_p]
        _lib.epi_mask_oracle_rect.restype = None
        _lib.plucker_oracle.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
        _lib.plucker_oracle.restype = None
    return _lib


def epipolar_mask(F: torch.Tensor, H: int, W: int, d: int) -> torch.Tensor:
    """F [B,T1,T2,3,3] fp32 -> bool [B, T1*H*W, T2*H*W]  (camcontexti2v.py:202-271; T1 = T2 in the UNet, T1 = 16 targets and
    T2 = 1 + n context frames for the adaptor's conditional mask, camcontexti2v.py:493-521)."""
    lib = _load()
    Fm = np.ascontiguousarray(F.detach().cpu().numpy().astype(np.float32))
    B, T1, T2 = Fm.shape[0], Fm.shape[1], Fm.shape[2]
    out = np.empty((B, T1 * H * W, T2 * H * W), dtype=np.uint8)
    lib.epi_mask_oracle_rect(Fm.ctypes.data, B, T1, T2, H, W, d, out.ctypes.data)
    return torch.from_numpy(out).bool()


def plucker(K: torch.Tensor, c2w: torch.Tensor, H: int, W: int, mode: str = "plucker") -> torch.Tensor:
    """K [B,T,3,3], c2w [B,T,4,4] -> [B,6,T,H,W] fp32 (base.py:112-174)."""
    lib = _load()
    Kn = np.ascontiguousarray(K.detach().cpu().numpy().astype(np.float32))
    Cn = np.ascontiguousarray(c2w.detach().cpu().numpy().astype(np.float32))
    B, T = Kn.shape[0], Kn.shape[1]
    out = np.empty((B, 6, T, H, W), dtype=np.float32)
    lib.plucker_oracle(Kn.ctypes.data, Cn.ctypes.data, B, T, H, W, 1 if mode == "plucker" else 0, out.ctypes.data)
    return torch.from_numpy(out)
