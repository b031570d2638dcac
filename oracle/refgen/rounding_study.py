"""Where does the 16-bit operand error of the CUDA path come from?  (TEST INFRASTRUCTURE, CPU only.)

Runs the fp32 oracle on the small model with emulated operand rounding at chosen places and prints the rel-L2 error of the
UNet pass against the unrounded oracle:  python oracle/refgen/rounding_study.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from camc2v_b200 import synth  # noqa: E402
from camc2v_b200.config import UNetConfig  # noqa: E402
from camc2v_b200.modules import build_unet  # noqa: E402
from camc2v_b200.testing import synth_unet_inputs  # noqa: E402
from oracle import camera_oracle, unet_oracle  # noqa: E402

_lin, _c2, _c3, _att = F.linear, F.conv2d, F.conv3d, unet_oracle.softmax_attention
MODE = dict(act=None, w=None, attn=None)


def rd(t, dt):
    return t if dt is None or t is None else t.to(dt).float()


def lin(x, w, b=None):
    return _lin(rd(x, MODE["act"]), rd(w, MODE["w"]), b)


def c2(x, w, b=None, **k):
    return _c2(rd(x, MODE["act"]), rd(w, MODE["w"]), b, **k)


def c3(x, w, b=None, **k):
    return _c3(rd(x, MODE["act"]), rd(w, MODE["w"]), b, **k)


def att(q, k, v, heads, mask=None, q_chunk=2048, fused=False):
    dt = MODE["attn"]
    if dt is None:
        return _att(q, k, v, heads, mask, q_chunk, fused)
    B, Lq, C = q.shape
    D = C // heads
    qh, kh, vh = (rd(t, dt).view(B, -1, heads, D).transpose(1, 2) for t in (q, k, v))
    sim = qh @ kh.transpose(-1, -2) * D ** -0.5
    if mask is not None:
        sim = sim.masked_fill(~mask[:, None], float("-inf"))
    p = sim.softmax(-1)
    m = sim.amax(-1, keepdim=True)
    e = rd(torch.exp(sim - m), dt)                       # the kernel rounds exp(s - max) to 16 bits, normalises in fp32
    out = (e @ vh) / torch.exp(sim - m).sum(-1, keepdim=True)
    return rd(out.transpose(1, 2).reshape(B, Lq, C), dt)


def main():
    F.linear, F.conv2d, F.conv3d, unet_oracle.softmax_attention = lin, c2, c3, att
    cfg = UNetConfig(model_channels=64, origin_h=128, origin_w=128)
    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in build_unet(cfg).state_dict().items()}
    sd = synth.synth_state_dict(shapes, 0)
    inp = synth_unet_inputs(cfg, 16, 2, "small")
    K, w2c = synth.synth_camera("pan_yaw", T=16, H=128, W=128)
    torch.manual_seed(123)
    Fm = camera_oracle.fundamental_matrices(K, camera_oracle.relative_c2w(w2c, torch.zeros(1, dtype=torch.long)))
    masks = {d: oracle.epipolar_mask(Fm, 128 // d, 128 // d, d) for d in (8, 16, 32, 64)}
    cam = {"pluker_embedding_features": inp["pluker"], "sample_locs_dict": masks, "add_type": "add_to_main_branch"}
    orc = unet_oracle.UNetOracle(sd, cfg)
    xc = torch.cat([inp["x"], inp["c_concat"]], dim=1)
    t = torch.full((1,), 599, dtype=torch.long)

    def run(**m):
        MODE.update(act=None, w=None, attn=None)
        MODE.update(m)
        return orc.forward(xc, t, inp["ctx_cond"], inp["fs"], cam).double()

    y0 = run()
    bf, hf = torch.bfloat16, torch.float16
    for name, m in [("weights bf16", dict(w=bf)), ("GEMM activations bf16", dict(act=bf)), ("attention q,k,v,P,out bf16", dict(attn=bf)),
                    ("all bf16 (the CUDA path)", dict(w=bf, act=bf, attn=bf)), ("activations+attention fp16, weights bf16", dict(w=bf, act=hf, attn=hf)),
                    ("all fp16", dict(w=hf, act=hf, attn=hf)), ("attention fp16, rest bf16", dict(w=bf, act=bf, attn=hf))]:
        y = run(**m)
        print(f"{name:45s} rel-L2 {float((y - y0).norm() / y0.norm()):.3e}   max-norm {float((y - y0).abs().max() / y0.abs().max()):.3e}")


if __name__ == "__main__":
    main()
