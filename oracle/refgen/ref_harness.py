"""Harness that imports the UNMODIFIED reference (LDenninger/CamC2V) from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  It is used (a) to pin `oracle/` against the reference's own code and
(b) to generate the golden vectors committed under `tests/golden/` (see `make_golden.py`).
Nothing in the product path (`camc2v_b200/`) imports it.  On the GPU box `/root/reference` does not exist: there the harness
imports the byte-identical copies that `install_ref.py` placed under `oracle/_ref/` (git-ignored, shipped with the snapshot), and
only `bench.py --impl reference`, `bench.py`'s `cpu_baseline` leg and the drop-in tests (tests/test_dropin_gpu.py) use it.

What it does:
  * installs a ~20-line stub `pytorch_lightning` (absent in this image) so that
    `model.camcontexti2v` -> `model.base` -> `lvdm.models.ddpm3d` import;
  * loads `configs/models/camcontexti2v_256.yaml` with `yaml.safe_load` into an attribute-dict;
  * replaces the three frozen third-party stages (VAE, CLIP text, CLIP image) by `torch.nn.Identity`
    and the pose encoder (needs `diffusers`, absent) by None;
  * constructs `model.camcontexti2v.CamContextI2V`, which applies the reference's own monkey patches
    (camcontexti2v.py:111-170), then adds `pluker_projection` exactly as camcontexti2v.py:151-156
    would have done had a pose encoder been configured;
  * overrides `DDIMSampler.register_buffer` (ddim.py:18-22 hard-codes `.to("cuda")`).
"""
from __future__ import annotations

import copy
import os
import sys
import types

import torch
import torch.nn as nn
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.abspath(os.path.join(_HERE, "..", "_ref"))          # oracle/_ref (install_ref.py)


def _find_reference() -> str:
    env = os.environ.get("CAMC2V_REFERENCE")
    if env:
        return env
    if os.path.isdir("/root/reference/CamContextI2V"):
        return "/root/reference"
    return _INSTALLED


REF_ROOT = _find_reference()
REF_PKG = os.path.join(REF_ROOT, "CamContextI2V")
REF_CFG = os.path.join(REF_ROOT, "configs")


class AttrDict(dict):
    """dict with attribute access + hasattr/setattr, enough for the reference's OmegaConf usage."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(o):
    if isinstance(o, dict):
        return AttrDict({k: to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [to_attr(v) for v in o]
    return o


def _install_pl_stub():
    if "pytorch_lightning" in sys.modules:
        return
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    class LightningDataModule:
        pass

    def seed_everything(seed, *a, **k):
        torch.manual_seed(seed)
        return seed

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.seed_everything = seed_everything
    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_only = lambda f: f
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util


def setup_reference_imports():
    if not os.path.isdir(REF_PKG):
        raise RuntimeError(f"reference not found at {REF_PKG}")
    _install_pl_stub()
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)


def load_model_config(name="models/camcontexti2v_256.yaml"):
    with open(os.path.join(REF_CFG, name)) as f:
        cfg = yaml.safe_load(f)
    return to_attr(cfg["model"])


IDENTITY = {"target": "torch.nn.Identity"}


def build_reference_model(unet_overrides: dict | None = None, seed: int = 0, unet_target: str | None = None):
    """Construct the reference CamContextI2V (UNet + patches + Epipolar + pluker_projection) on CPU.

    unet_overrides: optional overrides of unet_config.params (e.g. a small model_channels for fast tests).
    unet_target: optional replacement of `unet_config.target` (the YAML swap of INTEGRATION.md section 2, e.g.
        "camc2v_b200.modules.UNetModel"); everything else - the reference's constructor with its by-name forward re-binding and
        sub-module injection, LatentDiffusion.apply_model, DiffusionWrapper - stays the reference's own code.
    """
    setup_reference_imports()
    from model.camcontexti2v import CamContextI2V  # noqa

    mc = load_model_config()
    p = mc.params
    p.first_stage_config = to_attr(IDENTITY)
    p.cond_stage_config = to_attr(IDENTITY)
    p.img_cond_stage_config = to_attr(IDENTITY)
    p.image_proj_stage_config = to_attr(IDENTITY)
    p.pose_encoder_config = None
    # The 46.5 M adaptor is once-per-sample (SURVEY f-1) and not on the per-step path.
    p.multi_cond_strategy = None
    p.pop("multi_latent_adaptor", None)
    if unet_overrides:
        for k, v in unet_overrides.items():
            p.unet_config.params[k] = v
    if unet_target:
        p.unet_config.target = unet_target
    torch.manual_seed(seed)
    try:
        model = CamContextI2V(**p)
    except TypeError:
        raise
    model.eval()
    unet = model.model.diffusion_model
    # camcontexti2v.py:151-156: pluker_projection is only added when a pose encoder exists.
    for _name, _module in unet.named_modules():
        if _module.__class__.__name__ == "BasicTransformerBlock" and hasattr(_module, "epipolar"):
            c = _module.attn1.to_k.in_features
            if not hasattr(_module, "pluker_projection"):
                _module.add_module("pluker_projection", nn.Linear(c, c))
    return model


def build_baseline_model(kind: str, unet_overrides: dict | None = None, seed: int = 0):
    """Construct one of the reference's baselines (R/baseline/*): 'cameractrl', 'motionctrl', 'cami2v'."""
    setup_reference_imports()
    import importlib
    target, cfgname = {
        "cameractrl": ("baseline.cameractrl.cameractrl.CameraCtrl", "baseline/cameractrl_256.yaml"),
        "motionctrl": ("baseline.motionctrl.motionctrl.MotionCtrl", "baseline/motionctrl_256.yaml"),
        "cami2v": ("baseline.cami2v.cami2v.CamI2V", "baseline/cami2v_256.yaml"),
    }[kind]
    mc = load_model_config(cfgname)
    p = mc.params
    for k in ("first_stage_config", "cond_stage_config", "img_cond_stage_config", "image_proj_stage_config"):
        p[k] = to_attr(IDENTITY)
    if "pose_encoder_config" in p:
        p.pose_encoder_config = None
    if unet_overrides:
        for k, v in unet_overrides.items():
            p.unet_config.params[k] = v
    mod, cls = target.rsplit(".", 1)
    klass = getattr(importlib.import_module(mod), cls)
    torch.manual_seed(seed)
    model = klass(**p)
    model.eval()
    return model


def patch_ddim_for_cpu():
    setup_reference_imports()
    from lvdm.models.samplers.ddim import DDIMSampler

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    DDIMSampler.register_buffer = register_buffer
    return DDIMSampler
