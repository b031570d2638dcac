"""GPU probe: run the CUDA UNet against the committed golden outputs of the reference and print errors/timings.

    python tools/unet_probe.py [small] [full] [--timing]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from camc2v_b200 import synth  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402
from camc2v_b200.modules import build_unet  # noqa: E402
from camc2v_b200.testing import synth_unet_inputs  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


def run(name, cfg, hw, gold_file, timing):
    dev = "cuda"
    t0 = time.time()
    model = build_unet(cfg)
    synth.fill_module_(model, seed=0)
    model = model.to(dev)
    print(f"[{name}] model built+filled in {time.time() - t0:.1f}s", flush=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", gold_file))
    inp = synth_unet_inputs(cfg, hw, 2, name)
    Fm = torch.from_numpy(gold["F"]).to(dev)
    xc = torch.cat([inp["x"], inp["c_concat"]], dim=1).to(dev)
    t = torch.full((1,), 599, dtype=torch.long, device=dev)
    fs = inp["fs"].to(dev)
    pl = [p.to(dev) for p in inp["pluker"]]
    cam = {"pluker_embedding_features": pl, "epipolar_F": Fm, "add_type": "add_to_main_branch"}
    for key, ctx, c in (("y_cond", inp["ctx_cond"], cam), ("y_uncond", inp["ctx_uncond"], cam), ("y_nocam", inp["ctx_cond"], None)):
        if key not in gold.files:
            continue
        y = model(xc, t, context=ctx.to(dev), fs=fs, camera_condition=c)
        torch.cuda.synchronize()
        r = rel(y.cpu(), torch.from_numpy(gold[key]))
        print(f"[{name}] {key}: rel-L2 {r[0]:.3e}  max|err|/max|ref| {r[1]:.3e}  finite={bool(torch.isfinite(y).all())}", flush=True)
    if timing:
        from camc2v_b200 import _lib
        ctx = inp["ctx_cond"].to(dev)
        for _ in range(2):
            model(xc, t, context=ctx, fs=fs, camera_condition=cam)
        torch.cuda.synchronize()
        n0 = _lib.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(3):
            model(xc, t, context=ctx, fs=fs, camera_condition=cam)
        e1.record()
        torch.cuda.synchronize()
        print(f"[{name}] cond pass: {e0.elapsed_time(e1) / 3:.2f} ms GPU, {(time.time() - t0) / 3 * 1e3:.2f} ms wall, "
              f"{(_lib.LAUNCHES - n0) // 3} C-ABI calls/pass, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    timing = "--timing" in args
    which = [a for a in args if not a.startswith("--")] or ["small", "full"]
    if "small" in which:
        run("small", UNetConfig(model_channels=64, origin_h=128, origin_w=128), 16, "unet_small.npz", timing)
    if "full" in which:
        run("full", UNetConfig(), 32, "unet_full.npz", timing)
