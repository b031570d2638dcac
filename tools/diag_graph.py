"""Diagnostic: one full-size CFG step, eager vs CUDA-graph replay (serial / concurrent passes); prints the deviations.
    python tools/diag_graph.py            (C2V_PDL=0 in the environment switches programmatic dependent launch off)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from camc2v_b200 import synth  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402
from camc2v_b200.modules import build_unet  # noqa: E402
from camc2v_b200.sampler import DDIMSampler, DenoiserModel  # noqa: E402
from camc2v_b200.testing import synth_unet_inputs  # noqa: E402

DEV = "cuda"
small = "--small" in sys.argv
cfg = UNetConfig(model_channels=64, origin_h=128, origin_w=128) if small else UNetConfig()
unet = build_unet(cfg)
synth.fill_module_(unet, seed=0)
unet = unet.to(DEV)
g = np.load(os.path.join(ROOT, "tests", "golden", "unet_small.npz" if small else "unet_full.npz"))
inp = synth_unet_inputs(cfg, 16 if small else 32, 2, "small" if small else "full")
cam = {"pluker_embedding_features": [p.to(DEV) for p in inp["pluker"]], "epipolar_F": torch.from_numpy(g["F"]).to(DEV), "add_type": "add_to_main_branch"}
model = DenoiserModel(unet).to(DEV)


def conds():
    cond = {"c_crossattn": [inp["ctx_cond"].to(DEV)], "c_concat": [inp["c_concat"].to(DEV)], "camera_condition": cam}
    uc = {"c_crossattn": [inp["ctx_uncond"].to(DEV)], "c_concat": [inp["c_concat"].to(DEV)]}
    return cond, uc


x = inp["x"].to(DEV)
t = torch.full((1,), 999, dtype=torch.long, device=DEV)
kw = dict(fs=inp["fs"].to(DEV), enable_camera_condition=True)
cond, uc = conds()
uc["camera_condition"] = cam
ref = [model.apply_model(x, t, c, **kw).clone() for c in (cond, uc)]
ref2 = [model.apply_model(x, t, c, **kw).clone() for c in (cond, uc)]
print("eager run-to-run identical:", [bool(torch.equal(a, b)) for a, b in zip(ref, ref2)])
for conc in (False, True):
    for trial in range(2):
        s = DDIMSampler(model, concurrent_passes=conc)
        cond, uc = conds()
        uc["camera_condition"] = cam
        outs = s._unet_passes(x, t, [cond, uc], kw, True)
        outs2 = [o.clone() for o in s._unet_passes(x, t, [cond, uc], kw, True)]
        torch.cuda.synchronize()
        d = [float((o - r).abs().max() / r.abs().max()) for o, r in zip(outs2, ref)]
        print(f"graph concurrent={conc} trial {trial}: max-norm deviation from eager cond / uncond = {d[0]:.3e} / {d[1]:.3e}")
