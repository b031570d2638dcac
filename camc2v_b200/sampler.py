"""DDIM / classifier-free-guidance sampling loop on the CUDA path.

Mirrors the sampler <-> model contract of the reference:
    DDIMSampler.make_schedule / sample / ddim_sampling / p_sample_ddim   R/lvdm/models/samplers/ddim.py:24-346
    LatentDiffusion.apply_model + DiffusionWrapper.forward ('hybrid')     R/lvdm/models/ddpm3d.py:724-739, 1268-1272
    DDPM.register_schedule (linear betas)                                 R/lvdm/models/ddpm3d.py:125-188
The sampler works with any object that honours that contract (`apply_model`, `alphas_cumprod`, `betas`,
`num_timesteps`, `parameterization`, `use_dynamic_rescale`, `device`) — e.g. the reference's own
LightningModule whose UNet has been swapped for camc2v_b200.modules.UNetModel — or with
the light `DenoiserModel` below.

What changes versus the reference: the ~15 elementwise / reduction launches of the CFG combine, guidance
rescale and DDIM update are ONE kernel (ops.cfg_ddim_update); the schedule scalars stay on the host (no
`.item()`-style device syncs per step); the camera condition is not deep-copied every step (ddim.py:259);
and the two UNet passes of a step can be replayed from a CUDA graph (`use_cuda_graph=True`).
"""
from __future__ import annotations

from typing import Optional

import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .modules import UNetModel


_DEBUG_GRAPH = os.environ.get("C2V_DEBUG_GRAPH", "0") == "1"


def make_beta_schedule_linear(n_timestep=1000, linear_start=0.00085, linear_end=0.012) -> np.ndarray:
    return (np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2)


def make_ddim_timesteps(method: str, n_ddim: int, n_ddpm: int) -> np.ndarray:
    if method == "uniform":
        c = n_ddpm // n_ddim
        return np.asarray(list(range(0, n_ddpm, c))) + 1
    if method == "uniform_trailing":
        c = n_ddpm / n_ddim
        return np.flip(np.round(np.arange(n_ddpm, 0, -c))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(n_ddpm * .8), n_ddim)) ** 2).astype(int) + 1
    raise NotImplementedError(f'There is no ddim discretization method called "{method}"')


class DiffusionWrapper(nn.Module):
    """ddpm3d.py:1251-1318, conditioning_key == 'hybrid' (the DynamiCrafter / CamContextI2V setting)."""

    def __init__(self, diffusion_model: UNetModel, conditioning_key: str = "hybrid"):
        super().__init__()
        assert conditioning_key == "hybrid"
        self.diffusion_model = diffusion_model
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat: list = None, c_crossattn: list = None, **kwargs):
        xc = torch.cat([x] + c_concat, dim=1)
        # A single context tensor is passed through as it is: its address is then stable over the sampling loop, which is what
        # lets make_context_pack / CrossAttention._context_kv keep the bf16 tokens and the projected K/V per sample.
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(xc, t, context=cc, **kwargs)


class DenoiserModel(nn.Module):
    """The part of LatentDiffusion the sampler touches: schedule buffers + apply_model (ddpm3d.py:125-188, 724-739)."""

    def __init__(self, unet: UNetModel, timesteps=1000, linear_start=0.00085, linear_end=0.012, parameterization="eps"):
        super().__init__()
        assert parameterization == "eps"
        self.model = DiffusionWrapper(unet)
        self.parameterization = parameterization
        self.use_dynamic_rescale = False
        self.num_timesteps = timesteps
        betas = make_beta_schedule_linear(timesteps, linear_start, linear_end)
        ac = np.cumprod(1.0 - betas, axis=0)
        self.register_buffer("betas", torch.tensor(betas, dtype=torch.float32))
        self.register_buffer("alphas_cumprod", torch.tensor(ac, dtype=torch.float32))
        self.register_buffer("alphas_cumprod_prev", torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32))

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond, **kwargs):
        if not isinstance(cond, dict):
            raise NotImplementedError("hybrid conditioning expects a dict with c_concat / c_crossattn")
        out = self.model(x_noisy, t, **cond, **kwargs)
        return out[0] if isinstance(out, tuple) else out


def _leaf_key(o):
    """Identity of a conditioning structure for the CUDA-graph cache: every tensor leaf by (id, address, shape, dtype), plain
    values by value.  The graph entry also keeps the structure itself alive, so neither ids nor addresses can be recycled.
    In-place refills of a leaf (same address) deliberately do NOT change the key: that is the supported way to reuse a captured
    graph for the next video (refresh_camera_caches / refresh_context_caches re-derive the cached state in place)."""
    if isinstance(o, torch.Tensor):
        return ("T", id(o), o.data_ptr(), tuple(o.shape), str(o.dtype))
    if isinstance(o, dict):
        return ("D",) + tuple((k, _leaf_key(v)) for k, v in o.items())
    if isinstance(o, (list, tuple)):
        return ("L",) + tuple(_leaf_key(v) for v in o)
    if o is None or isinstance(o, (bool, int, float, str)):
        return ("V", o)
    return ("O", id(o))


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        if getattr(model, "parameterization", "eps") != "eps":
            raise NotImplementedError(f"parameterization {model.parameterization!r}: only eps-prediction models are sampled (ddim.py:285-288)")
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self._graph = None
        self.concurrent_passes = kwargs.get("concurrent_passes", True)
        self.cfg_pair = kwargs.get("cfg_pair", None)       # parallel.CfgPair: split the CFG halves over two ranks

    def reset_graph(self):
        """Drop the captured CUDA graph (and the references it holds to conditioning tensors and cache-derived buffers)."""
        if self._graph is not None:
            from . import modules
            modules.unpin_caches(self._graph.get("pin"))
        self._graph = None

    def __del__(self):
        try:
            self.reset_graph()          # a dead sampler must not leave its cache entries pinned forever
        except Exception:
            pass

    # -------------------------------------------------------------------------------------------- schedule
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """ddim.py:24-57; all per-step coefficients are kept as host fp32 numpy arrays."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps)
        ac = self.model.alphas_cumprod.detach().float().cpu()
        assert ac.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        if getattr(self.model, "use_dynamic_rescale", False):
            raise NotImplementedError("use_dynamic_rescale")
        acn = ac.numpy()
        a = acn[self.ddim_timesteps].astype(np.float64)
        a_prev = np.asarray([acn[0]] + acn[self.ddim_timesteps[:-1]].tolist(), dtype=np.float64)
        sig = ddim_eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
        self.ddim_alphas = a.astype(np.float32)
        self.ddim_alphas_prev = a_prev.astype(np.float32)
        self.ddim_sigmas = sig.astype(np.float32)
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1. - a).astype(np.float32)

    # -------------------------------------------------------------------------------------------- sampling
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None, quantize_x0=False,
               eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=False,
               schedule_verbose=False, x_T=None, log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None,
               precision=None, fs=None, timestep_spacing='uniform', guidance_rescale=0.0, use_cuda_graph=False, **kwargs):
        """ddim.py:59-132, same keyword list.  Options outside the CamContextI2V sampling protocol raise NotImplementedError
        (they are never silently ignored): quantize_x0, mask / x0 blending, noise_dropout, score_corrector, precision=16."""
        if quantize_x0 or mask is not None or x0 is not None or noise_dropout > 0. or score_corrector is not None or precision is not None:
            raise NotImplementedError("DDIMSampler.sample: quantize_x0 / mask / x0 / noise_dropout / score_corrector / precision are "
                                      "outside the per-step scope (ddim.py:176-183, 289-296, 340-343)")
        self.make_schedule(ddim_num_steps=S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=False)
        size = (batch_size,) + tuple(shape)
        return self.ddim_sampling(conditioning, size, x_T=x_T, temperature=temperature, callback=callback, img_callback=img_callback,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, fs=fs,
                                  guidance_rescale=guidance_rescale, use_cuda_graph=use_cuda_graph, **kwargs)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, fs=None, guidance_rescale=0.0, log_every_t=100, use_cuda_graph=False,
                      img_callback=None, callback=None, **kwargs):
        """ddim.py:134-238 (the mask / paste / noise-shaping editing branches are outside the per-step scope)."""
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        timesteps = self.ddim_timesteps
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        total_steps = timesteps.shape[0]
        for i, step in enumerate(np.flip(timesteps)):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, temperature=temperature,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning, fs=fs,
                                              guidance_rescale=guidance_rescale, use_cuda_graph=use_cuda_graph, **kwargs)
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    # -------------------------------------------------------------------------------------------- one step
    def _unet_passes(self, x, t, conds, kwargs, use_cuda_graph):
        """apply_model for every conditioning in `conds` (the cond / uncond passes of a CFG step, ddim.py:262-263, or just
        one of them under CFG-split), optionally replayed from one CUDA graph with the passes as parallel branches."""
        if not use_cuda_graph:
            return [self.model.apply_model(x, t, c, **kwargs) for c in conds]
        g = self._graph
        key = (_leaf_key(conds), _leaf_key(kwargs), tuple(x.shape), str(x.dtype))
        if g is None or g["key"] != key:
            self.reset_graph()
            from . import modules
            sx, st = x.clone(), t.clone()
            # Every per-sample cache entry the passes touch (context packs, projected K/V, tile maps, packed masks, channels-last
            # Pluecker copies) is recorded from the warm-up on: recorded entries cannot be evicted, so the capture below only HITS
            # entries the warm-up created outside the graph, and exactly these entries are pinned for the lifetime of the graph.
            modules.begin_cache_record()
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):          # warm-up outside capture: builds weight packs, fills allocator pools
                    for c in conds:
                        self.model.apply_model(sx, st, c, **kwargs)
                torch.cuda.current_stream().wait_stream(side)
                inserts0 = modules.CACHE_INSERTS
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    # The cond and uncond passes are independent: capture them as parallel branches of the graph so that
                    # the small grids of the 8x8 / 4x4 levels and every kernel's last partial wave overlap with the other pass.
                    main = torch.cuda.current_stream()
                    branches = []
                    if self.concurrent_passes:
                        for _ in conds[1:]:               # fork BEFORE anything is captured on main: the branches have no
                            br = torch.cuda.Stream()      # dependency on the first pass
                            br.wait_stream(main)
                            branches.append(br)
                    outs = [self.model.apply_model(sx, st, conds[0], **kwargs)]
                    for i, c in enumerate(conds[1:]):
                        if self.concurrent_passes:
                            with torch.cuda.stream(branches[i]):
                                outs.append(self.model.apply_model(sx, st, c, **kwargs))
                        else:
                            outs.append(self.model.apply_model(sx, st, c, **kwargs))
                    for br in branches:
                        main.wait_stream(br)
            finally:
                touched = modules.end_cache_record()
            if modules.CACHE_INSERTS != inserts0:
                # a derived buffer was produced INSIDE the capture: with parallel branches another branch may read it without a
                # dependency edge - never replay such a graph
                raise RuntimeError("camc2v_b200.sampler: a per-sample cache entry was created during CUDA-graph capture "
                                   "(the warm-up pass should have created it); refusing to replay a graph with unordered producers")
            # `refs` keeps every conditioning tensor of the key alive; `pin` marks the cache-derived buffers the captured kernels
            # point at as non-evictable.
            g = self._graph = dict(key=key, graph=graph, x=sx, t=st, outs=outs, ec=outs[0], eu=outs[-1], refs=(list(conds), dict(kwargs)),
                                   pin=modules.pin_entries(touched))
        g["x"].copy_(x)
        g["t"].copy_(t)
        g["graph"].replay()
        if _DEBUG_GRAPH:          # C2V_DEBUG_GRAPH=1: every replay is checked against an eager run of the same passes
            eager = [self.model.apply_model(x, t, c, **kwargs) for c in conds]
            dev = [float((o - e).abs().max() / e.abs().max()) for o, e in zip(g["outs"], eager)]
            if max(dev) > 0:
                print(f"[c2v debug] graph replay deviates from eager: {dev}", flush=True)
        return g["outs"]

    def _unet_pair(self, x, t, c, uc, kwargs, use_cuda_graph):
        ec, eu = self._unet_passes(x, t, [c, uc], kwargs, use_cuda_graph)
        return ec, eu

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, uc_type=None,
                      conditional_guidance_scale_temporal=None, mask=None, x0=None, guidance_rescale=0.0, noise=None,
                      use_cuda_graph=False, **kwargs):
        """ddim.py:241-346.  `noise` (optional) lets a caller supply the eta-noise; otherwise torch.randn is drawn at
        the same point of the RNG stream as the reference (ddim.py:340)."""
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0. or repeat_noise:
            raise NotImplementedError("option outside the CamContextI2V sampling protocol")
        if mask is not None or x0 is not None:
            raise NotImplementedError("mask / x0 blending (ddim.py:176-183) is outside the per-step scope")
        for opt in ("paste_cond_frame", "paste_overlap_frames", "noise_shaping", "timesteps", "precision", "clean_cond"):
            if kwargs.get(opt):
                raise NotImplementedError(f"{opt} (ddim.py:150-238 editing branches) is outside the per-step scope")
            kwargs.pop(opt, None)
        if getattr(self.model, "parameterization", "eps") != "eps":
            raise NotImplementedError("only eps-prediction models are sampled (ddim.py:285-288)")
        a_t = float(self.ddim_alphas[index])
        a_prev = float(self.ddim_alphas_prev[index])
        sigma_t = float(self.ddim_sigmas[index])
        s1m = float(self.ddim_sqrt_one_minus_alphas[index])
        camera_cfg = float(kwargs.pop("camera_cfg", 1.0))
        camera_cfg_scheduler = kwargs.pop("camera_cfg_scheduler", "constant")
        if camera_cfg_scheduler not in ("constant", "cosine"):
            raise NotImplementedError(camera_cfg_scheduler)
        e_nc, cam_w = None, 0.0
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_c = self.model.apply_model(x, t, c, **kwargs)
            e_u, scale, phi = e_c, 1.0, 0.0
        else:
            uc = unconditional_conditioning
            if kwargs.get("enable_camera_condition", False) and isinstance(c, dict):
                # the reference deep-copies the masks into uc every step (ddim.py:259-260); sharing the dict is equivalent
                cam = c.get("camera_condition")
                if cam is not None and uc.get("camera_condition") is not cam:
                    uc["camera_condition"] = cam
            if camera_cfg != 1.0 and kwargs.get("enable_camera_condition", False) and isinstance(c, dict):
                # camera guidance (ddim.py:268-280): a third pass, conditional but without the camera condition
                if self.cfg_pair is not None:
                    raise NotImplementedError("camera_cfg != 1 together with CFG-split")
                # a shallow copy per step; the CUDA-graph key compares tensor leaves, not dict identities
                c_nc = {k: v for k, v in c.items() if k != "camera_condition"}
                e_c, e_u, e_nc = self._unet_passes(x, t, [c, uc, c_nc], kwargs, use_cuda_graph)
                # scheduler weight from the HOST copy of the timestep (ddim_timesteps[index] == t): no device sync
                w = 1.0 if camera_cfg_scheduler == "constant" else math.cos((1.0 - float(self.ddim_timesteps[index]) / 999.0) * math.pi / 2.0)
                cam_w = (camera_cfg - 1.0) * w
            elif self.cfg_pair is not None:
                # CFG-split (parallel.CfgPair): this rank runs ONE of the two passes, the pair exchanges the predictions
                e_loc = self._unet_passes(x, t, [c if self.cfg_pair.role == 0 else uc], kwargs, use_cuda_graph)[0]
                e_c, e_u = self.cfg_pair.exchange(e_loc)
                if noise is None:
                    noise = self.cfg_pair.noise(x.shape, x.device)
            else:
                e_c, e_u = self._unet_pair(x, t, c, uc, kwargs, use_cuda_graph)
            scale, phi = float(unconditional_guidance_scale), float(guidance_rescale)
        if noise is None:
            noise = torch.randn(x.shape, device=x.device)
        if temperature != 1.:
            noise = noise * temperature
        return ops.cfg_ddim_update(x.float(), e_c, e_u, noise, scale, phi, a_t, a_prev, sigma_t, s1m, e_cond_nocam=e_nc, cam_weight=cam_w)
