"""GPU parity of the once-per-sample MultiLatentEpipolarAdaptor against the golden outputs of the
unmodified reference classes (tests/golden/adaptor_small.npz: adaptors.py:36-182 + the conditional mask of
camcontexti2v.py:493-521).  Mask bit-exact; adaptor output within the tolerance of tests/test_unet_gpu.py."""
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


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLD, "adaptor_small.npz"))
    return g, json.loads(str(g["kwargs"]))


def test_conditional_mask_bit_exact(gold):
    from camc2v_b200 import camera, ops
    g, _ = gold
    K, w2c, w2c_cond = (torch.from_numpy(g[k]) for k in ("K", "w2c", "w2c_cond"))
    Fm = camera.conditional_fundamental_matrices(K, w2c, w2c_cond, torch.zeros(1, dtype=torch.long))
    assert np.allclose(Fm.numpy(), g["F"], rtol=1e-5, atol=1e-7)
    m = ops.epipolar_mask(torch.from_numpy(g["F"]).to(DEV), 8, 8, 8)                  # rectangular: 16 target x 3 context frames
    assert m.shape == (1, 1024, 192)
    assert np.array_equal(np.packbits(m.cpu().numpy(), axis=-1), g["mask_packed"])


def test_adaptor_vs_reference_golden(gold):
    from camc2v_b200 import ops, synth
    from camc2v_b200.adaptor import MultiLatentEpipolarAdaptor
    g, kw = gold
    m = MultiLatentEpipolarAdaptor(**kw)
    synth.fill_module_(m, seed=5)
    m = m.to(DEV)
    mask = ops.epipolar_mask(torch.from_numpy(g["F"]).to(DEV), 8, 8, 8)
    z = synth.synth_tensor("adaptor.z", (1, 192, 4), 9).to(DEV)
    for key, mk in (("y", mask), ("y_nomask", None)):
        y = m(z, mk)
        assert y.shape == (1, 1024, 4) and torch.isfinite(y).all()
        l2, mx = rel(y, torch.from_numpy(g[key]))
        print(f"adaptor {key}: rel-L2 {l2:.3e} max-norm {mx:.3e}")
        # the last op is a LayerNorm over only output_dim = 4 channels, which amplifies the worst element: 2x bound on the max-norm
        assert l2 < TOL and mx < 2 * TOL, (key, l2, mx)
    # batch of two = two independent samples
    y2 = m(torch.cat([z, z.flip(1)], 0), torch.cat([mask, mask], 0))
    assert rel(y2[0], m(z, mask)[0])[0] < TOL


def test_resampler_vs_reference_golden():
    """Resampler image-token projector (SURVEY f-4, resampler.py:100-166) against the unmodified reference class."""
    from camc2v_b200 import synth
    from camc2v_b200.resampler import Resampler
    g = np.load(os.path.join(GOLD, "resampler_small.npz"))
    kw = json.loads(str(g["kwargs"]))
    m = Resampler(**kw)
    synth.fill_module_(m, seed=6)
    m = m.to(DEV)
    x = synth.synth_tensor("resampler.x", (2, 33, 96), 10).to(DEV)
    y = m(x)
    assert y.shape == (2, 64, 128) and torch.isfinite(y).all()
    l2, mx = rel(y, torch.from_numpy(g["y"]))
    print(f"resampler: rel-L2 {l2:.3e} max-norm {mx:.3e}")
    assert l2 < TOL and mx < 2 * TOL, (l2, mx)


def test_vae_decoder_vs_reference_golden():
    """decode_first_stage (SURVEY f-3): AutoencoderKL.decode / ae_modules.Decoder against the unmodified reference classes."""
    from camc2v_b200 import synth
    from camc2v_b200.vae import AutoencoderKLDecoder
    g = np.load(os.path.join(GOLD, "vae_small.npz"))
    dd = json.loads(str(g["ddconfig"]))
    m = AutoencoderKLDecoder(dd)
    synth.fill_module_(m, seed=7)
    m = m.to(DEV)
    z = synth.synth_tensor("vae.z", (2, 4, 8, 8), 11).to(DEV)
    y = m.decode(z)
    assert y.shape == (2, 3, 64, 64) and torch.isfinite(y).all()
    l2, mx = rel(y, torch.from_numpy(g["y"]))
    print(f"vae decoder: rel-L2 {l2:.3e} max-norm {mx:.3e}")
    assert l2 < TOL and mx < 2 * TOL, (l2, mx)


def test_conv3x3_on_images_wider_than_a_tile():
    """256-pixel rows (the VAE decoder's last level): one 128-pixel tile is half an image row; TMA zero fill is the padding."""
    from camc2v_b200 import ops
    NB, H, W, Cin, Cout = 2, 8, 256, 64, 128
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(NB, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) * (9 * Cin) ** -0.5
    b = torch.randn(Cout, generator=gen)
    xb, wb = x.to(ops.BF16), w.to(ops.BF16)
    ref = torch.nn.functional.conv2d(xb.float(), wb.float(), b, padding=1).permute(0, 2, 3, 1).reshape(NB * H * W, Cout)
    a = xb.permute(0, 2, 3, 1).reshape(NB * H * W, Cin).contiguous().to(DEV)
    wk = wb.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous().to(DEV)
    out = ops.conv3x3(a, wk, NB, H, W, bias=b.to(DEV))
    assert (out.cpu() - ref).abs().max() <= 2e-3 * ref.abs().max()


def test_vae_encoder_vs_reference_golden():
    """encode_first_stage up to the posterior moments (SURVEY f-3): ae_modules.Encoder + quant_conv against the reference classes."""
    from camc2v_b200 import synth
    from camc2v_b200.vae import AutoencoderKLEncoder
    g = np.load(os.path.join(GOLD, "vae_enc_small.npz"))
    dd = json.loads(str(g["ddconfig"]))
    m = AutoencoderKLEncoder(dd)
    synth.fill_module_(m, seed=8)
    m = m.to(DEV)
    x = synth.synth_tensor("vae.x", (2, 3, 64, 64), 12).to(DEV)
    mom = m.encode(x)
    assert mom.shape == (2, 8, 8, 8) and torch.isfinite(mom).all()
    l2, mx = rel(mom, torch.from_numpy(g["moments"]))
    print(f"vae encoder: rel-L2 {l2:.3e} max-norm {mx:.3e}")
    assert l2 < TOL and mx < 2 * TOL, (l2, mx)
