"""GPU parity of the once-per-sample CameraPoseEncoder against tests/golden/pose_encoder_small.npz - the
reference's own class (camera_pose_encoder.py:295-376) run on the restated diffusers stand-ins, see oracle/pose_encoder_oracle.py -
and of its four dedicated kernels against plain torch fp32 restatements of the same ops."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"
TOL = float(os.environ.get("C2V_TEST_TOL", "5e-3"))


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("npz,seed", [("pose_encoder_small.npz", 8), ("pose_encoder_generic.npz", 9)])
def test_pose_encoder_vs_reference_golden(npz, seed):
    from camc2v_b200.pose_encoder import CameraPoseEncoder
    from test_oracle_golden import _pose_inputs, pose_state_dict
    g = np.load(os.path.join(GOLD, npz))
    kw = json.loads(str(g["kwargs"]))
    _, sd = pose_state_dict(kw, seed)
    m = CameraPoseEncoder(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    x = _pose_inputs().to(DEV)
    feats = m(x)
    assert len(feats) == len(kw["channels"])
    for i, f in enumerate(feats):
        ref = torch.from_numpy(g[f"f{i}"])
        assert f.shape == ref.shape and torch.isfinite(f).all()
        l2, mx = rel(f, ref)
        print(f"pose encoder level {i}: rel-L2 {l2:.3e} max-norm {mx:.3e}")
        assert l2 < TOL and mx < 2 * TOL, (i, l2, mx)
    # a batch of two = two independent samples
    f2 = m(torch.cat([x, x.flip(2)], 0))
    assert torch.equal(f2[0][:16], feats[0]) and f2[1].shape[0] == 32


@pytest.mark.parametrize("D,T,HW,heads", [(40, 16, 64, 8), (80, 16, 16, 8), (160, 16, 16, 8), (64, 8, 32, 5), (48, 5, 7, 3)])
def test_attention_temporal_hd(D, T, HW, heads):
    from camc2v_b200 import ops
    B = 2
    torch.manual_seed(D + T)
    qkv = torch.randn(B * T * HW, 3 * heads * D, device=DEV).to(ops.BF16)
    out = ops.attention_temporal_hd(qkv, B, T, HW, heads, D)
    q, k, v = (t.float().view(B, T, HW, heads, D).permute(0, 2, 3, 1, 4) for t in qkv.chunk(3, dim=1))      # [B, HW, heads, T, D]
    ref = torch.softmax(q @ k.transpose(-1, -2) / D ** 0.5, dim=-1) @ v
    ref = ref.permute(0, 3, 1, 2, 4).reshape(B * T * HW, heads * D)
    l2, mx = rel(out.float(), ref)
    assert l2 < (2e-3 if ops.BF16 == torch.float16 else 8e-3), (l2, mx)
    if D == 64:
        assert rel(out.float(), ops.attention_temporal(qkv, B, T, HW, heads).float())[0] < 4e-3


def test_pixel_unshuffle_avgpool_relu():
    from camc2v_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(2, 6, 3, 32, 48, device=DEV)
    got = ops.pixel_unshuffle_cl(x, 8)                                                     # rows (b, t, y, x)
    ref = torch.nn.functional.pixel_unshuffle(x.permute(0, 2, 1, 3, 4).reshape(6, 6, 32, 48), 8)     # [(b t), 384, 4, 6]
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 384)
    assert torch.equal(got, ref.to(ops.BF16))
    a = torch.randn(3 * 8 * 12, 64, device=DEV)
    p32, p16 = ops.avgpool2_cl(a, 3, 8, 12)
    refp = torch.nn.functional.avg_pool2d(a.view(3, 8, 12, 64).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 64)
    assert torch.allclose(p32, refp, rtol=1e-6, atol=1e-7) and torch.equal(p16, p32.to(ops.BF16))
    h = torch.randn(1000, 64, device=DEV).to(ops.BF16)
    assert torch.equal(ops.relu_(h.clone()), torch.relu(h))
