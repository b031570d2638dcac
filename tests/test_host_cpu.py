"""CPU tests of the host-side logic: C-ABI surface, module topology / state_dict contract, schedule, camera glue,
FLOP model, rank sharding over gloo (world_size 2).  No CUDA kernel is launched."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------------------ C ABI
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "camc2v_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(c2v_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    from camc2v_b200 import _lib, build
    path = build.build()
    path16 = build.build(fp16=True)               # both operand flavours ship: bf16 and (default) IEEE half
    assert os.path.exists(path) and os.path.exists(path16)
    assert ctypes.CDLL(path).c2v_operand_dtype() == 0 and ctypes.CDLL(path16).c2v_operand_dtype() == 1
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/camc2v_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == declared, "camc2v_b200/_lib.py prototypes out of sync with the header"
    loaded = _lib.load()
    assert loaded.c2v_abi_version() == 1
    assert loaded.c2v_status_string(4) == b"unsupported shape"
    assert loaded.c2v_gemm_tile_n(320, 0) == 160 and loaded.c2v_gemm_tile_n(512, 0) == 128 and loaded.c2v_gemm_tile_n(4, 0) == 64
    assert loaded.c2v_gemm_tile_n(2560, 1) == 256 and loaded.c2v_gemm_tile_n(4096, 1) == 256 and loaded.c2v_gemm_tile_n(10240, 1) == 160


def test_structs_match_header_layout():
    """ctypes mirrors of c2v_gemm_desc / c2v_attn_desc: compile a tiny C program against the header and compare sizeof/offsetof."""
    from camc2v_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "camc2v_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(c2v_gemm_desc), offsetof(c2v_gemm_desc, M), offsetof(c2v_gemm_desc, epi), offsetof(c2v_gemm_desc, lda), offsetof(c2v_gemm_desc, out));
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(c2v_attn_desc), offsetof(c2v_attn_desc, kv_div), offsetof(c2v_attn_desc, k2), offsetof(c2v_attn_desc, epi_F), offsetof(c2v_attn_desc, mask), offsetof(c2v_attn_desc, mask_bstride));
  return 0;
}'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    g = [int(v) for v in out[0].split()]
    a = [int(v) for v in out[1].split()]
    G, A = _lib.GemmDesc, _lib.AttnDesc
    assert g == [ctypes.sizeof(G), G.M.offset, G.epi.offset, G.lda.offset, G.out.offset]
    assert a == [ctypes.sizeof(A), A.kv_div.offset, A.k2.offset, A.epi_F.offset, A.mask.offset, A.mask_bstride.offset]


def test_missing_library_fails_loudly(monkeypatch):
    from camc2v_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcamc2v_b200.so")
    with pytest.raises(_lib.C2VError):
        _lib.load()


def test_ops_reject_cpu_tensors():
    from camc2v_b200 import _lib, ops
    with pytest.raises(_lib.C2VError):
        ops.linear(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    with pytest.raises(_lib.C2VError):
        ops.groupnorm(torch.zeros(128, 64), torch.ones(64), torch.zeros(64), 1, 128, 1e-5, True)


def test_persistent_gemm_plan():
    """Host logic of gemm_ps.cu (gemm_ps_plan): the multi-wave 16-bit-output projections of the 32x32 / 16x16 levels go to the
    persistent kernel - with the weight tile resident in shared memory where its whole K extent fits next to an A ring - and
    everything with a residual, an fp32 result or fewer than 300 output tiles stays on the one-tile-per-CTA kernel."""
    import ctypes as C
    from camc2v_b200 import _lib
    lib = _lib.load()

    def plan(M, N, K, epi=0, bf16=1, res=0):
        p = (C.c_int * 3)()
        assert lib.c2v_gemm_persistent_plan(M, N, K, epi, bf16, res, p) == 0
        return tuple(p)

    GEGLU = _lib.EPI_GEGLU
    assert plan(16384, 2560, 320, GEGLU) == (2, 256, 14)        # 160 KB weight tile + 4-stage A ring; 10 N tiles x 14 CTAs
    assert plan(16384, 960, 320) == (2, 240, 37)                # q|k|v, 32x32 level: four 240-wide N tiles, 37 CTAs each
    assert plan(4096, 1920, 640) == (2, 128, 9)                 # q|k|v, 16x16 level: K = 640 fits only a 128-wide tile
    m, bn, _ = plan(4096, 5120, 640, GEGLU)                     # GEGLU weights are interleaved per 160-wide tile: 200 KB does not fit
    assert (m, bn) == (1, 160)
    assert plan(1024, 10240, 1280, GEGLU)[:2] == (1, 160)
    for mode_bn_p in (plan(16384, 320, 320, bf16=0, res=1), plan(16384, 960, 320, bf16=0), plan(1024, 3840, 1280), plan(256, 1280, 1280)):
        assert mode_bn_p[0] == 0                                 # residual / fp32 out / < 300 tiles
    for M, N, K, epi in [(16384, 2560, 320, GEGLU), (16384, 960, 320, 0), (4096, 1920, 640, 0), (16384, 1536, 512, 0)]:
        mode, bn, P = plan(M, N, K, epi)
        if mode == 2:
            assert N % bn == 0 and P * (N // bn) <= 148 and (K // 64) * bn * 128 + 3 * 16384 <= 226 * 1024


# ------------------------------------------------------------------------------------------------ module contract
@pytest.mark.parametrize("name,kw", [("small", dict(model_channels=64, origin_h=128, origin_w=128)), ("full", {})])
def test_state_dict_keys_and_shapes_match_reference(name, kw):
    from camc2v_b200.config import UNetConfig
    from camc2v_b200.modules import build_unet
    with torch.device("meta"):
        m = build_unet(UNetConfig(**kw))
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    ref = json.load(open(os.path.join(GOLD, f"state_dict_{name}.json")))
    assert mine == ref
    n = sum(int(np.prod(s)) for s in ref.values())
    if name == "full":
        assert abs(n / 1e6 - 1500.9) < 0.1        # 1500.9 M parameters


def test_temporal_block_fused_projection_pack(monkeypatch):
    """Host logic of the fused output projection of the camera-conditioned temporal block (modified_forwards.py:519-533):
    x + pluker_projection(n + p) + attn1(n).to_out + Epipolar(n + p).to_out == [n + p | attn1 heads | epipolar heads] @ w_cat^T + b_cat + x."""
    from camc2v_b200 import modules, synth
    from camc2v_b200.config import UNetConfig
    from camc2v_b200.modules import build_unet
    unet = build_unet(UNetConfig(model_channels=64, origin_h=128, origin_w=128))
    synth.fill_module_(unet, seed=0)
    blocks = [m for m in unet.modules() if isinstance(m, modules.BasicTransformerBlock) and hasattr(m, "epipolar")]
    assert len(blocks) == 16                                      # 16 temporal blocks carry epipolar + pluker_projection
    blk = blocks[3]
    p = blk._prepare()
    pp, a1, ep = blk.pluker_projection, blk.attn1.to_out[0], blk.epipolar.epipolar_attn.to_out[0]
    C = pp.weight.shape[0]
    assert tuple(p["w_cat"].shape) == (C, 3 * C) and p["w_cat"].dtype == modules.BF16 and p["w_cat"].is_contiguous()
    for i, lin in enumerate((pp, a1, ep)):
        assert torch.equal(p["w_cat"][:, i * C:(i + 1) * C], lin.weight.detach().to(modules.BF16))
    assert torch.allclose(p["b_cat"], (pp.bias + a1.bias + ep.bias).detach().float())
    g = torch.Generator().manual_seed(0)
    cat = torch.randn(40, 3 * C, generator=g)
    three = sum(torch.nn.functional.linear(cat[:, i * C:(i + 1) * C], lin.weight.float(), lin.bias.float()) for i, lin in enumerate((pp, a1, ep)))
    one = cat @ torch.cat([pp.weight, a1.weight, ep.weight], dim=1).float().t() + p["b_cat"]
    assert torch.allclose(one, three, atol=1e-5)
    # parameters of a CHILD module rewritten in place change the key the forward compares
    key = lambda: tuple((id(t), t._version) for l in blk._fused_out_sources() for t in (l.weight, l.bias))
    assert p["cat_ver"] == key()
    with torch.no_grad():
        a1.weight.mul_(2.0)
    assert p["cat_ver"] != key()
    # blocks without the injected camera modules, the other variants and the A/B switch keep the three-GEMM form
    plain = [m for m in unet.modules() if isinstance(m, modules.BasicTransformerBlock) and not hasattr(m, "epipolar")]
    assert plain and all("w_cat" not in m._prepare() for m in plain[:2])
    monkeypatch.setattr(modules, "FUSE_TEMPORAL_OUT", False)
    assert "w_cat" not in blk._prepare()


def test_topology_matches_reference_ds_lists():
    from camc2v_b200.config import UNetConfig, build_topology
    topo = build_topology(UNetConfig())
    assert [b.ds for b in topo.input_blocks] == [1, 1, 1, 1, 2, 2, 2, 4, 4, 4, 8, 8]
    assert [b.ds for b in topo.output_blocks] == [8, 8, 8, 4, 4, 4, 2, 2, 2, 1, 1, 1]
    kinds = [[l.kind for l in b.layers] for b in topo.output_blocks]
    assert kinds[2] == ["res", "up"] and kinds[5] == ["res", "spatial", "temporal", "up"] and kinds[11] == ["res", "spatial", "temporal"]
    n_epi = sum(l.epipolar for b in topo.input_blocks + [topo.middle] + topo.output_blocks for l in b.layers)
    assert n_epi == 16 and not topo.init_attn.epipolar


def test_variant_modules_attach_like_the_reference():
    from camc2v_b200.config import UNetConfig
    from camc2v_b200.modules import build_unet
    with torch.device("meta"):
        cc = build_unet(UNetConfig(model_channels=64), variant="cameractrl")
        mc = build_unet(UNetConfig(model_channels=64), variant="motionctrl")
        none = build_unet(UNetConfig(model_channels=64), variant="none")
    k_cc = [k for k in cc.state_dict() if "cc_projection" in k]
    k_mc = [k for k in mc.state_dict() if "cc_projection" in k]
    assert len(k_cc) == 32 and len(k_mc) == 34                     # 16 blocks (+ init_attn for MotionCtrl) x (weight, bias)
    assert mc.state_dict()["init_attn.0.transformer_blocks.0.cc_projection.weight"].shape == (512, 524)
    assert not any("epipolar" in k or "pluker" in k for k in none.state_dict())


def test_geglu_interleave_is_a_permutation():
    from camc2v_b200 import ops
    w = torch.arange(2560, dtype=torch.float32)[:, None].repeat(1, 2)
    b = torch.arange(2560, dtype=torch.float32)
    w2, b2 = ops.geglu_interleave(w, b)
    assert sorted(b2.tolist()) == b.tolist()
    half = ops.tile_n(2560, ops.EPI_GEGLU) // 2          # value / gate columns per N tile (128 for the 256-wide tile)
    assert b2[:half].tolist() == list(range(half))       # tile 0 = value[0:half] | gate[0:half]
    assert b2[half:2 * half].tolist() == list(range(1280, 1280 + half))
    assert b2[2 * half:3 * half].tolist() == list(range(half, 2 * half))


# ------------------------------------------------------------------------------------------------ schedule / camera / flops
def test_sampler_schedule_matches_reference():
    from camc2v_b200.sampler import DDIMSampler, DenoiserModel
    g = np.load(os.path.join(GOLD, "unet_small.npz"))

    class Dummy(torch.nn.Module):
        pass
    m = DenoiserModel.__new__(DenoiserModel)
    torch.nn.Module.__init__(m)
    from camc2v_b200.sampler import make_beta_schedule_linear
    ac = np.cumprod(1.0 - make_beta_schedule_linear(), axis=0)
    m.register_buffer("alphas_cumprod", torch.tensor(ac, dtype=torch.float32))
    m.register_buffer("betas", torch.zeros(1000))
    m.num_timesteps = 1000
    m.use_dynamic_rescale = False
    s = DDIMSampler(m)
    s.make_schedule(25, "uniform_trailing", 1.0, verbose=False)
    assert np.array_equal(s.ddim_timesteps.astype(np.float64), g["sched.ddim_timesteps"])
    assert np.allclose(m.alphas_cumprod.numpy(), g["sched.alphas_cumprod"], rtol=2e-7)
    for mine, gk in ((s.ddim_alphas, "ddim_alphas"), (s.ddim_alphas_prev, "ddim_alphas_prev"), (s.ddim_sigmas, "ddim_sigmas"),
                     (s.ddim_sqrt_one_minus_alphas, "ddim_sqrt_one_minus_alphas")):
        assert np.allclose(mine, g["sched." + gk], rtol=2e-7, atol=0), gk
    s.make_schedule(50, "uniform", 0.0, verbose=False)
    assert s.ddim_timesteps[0] == 1 and len(s.ddim_timesteps) == 50 and float(s.ddim_sigmas.max()) == 0.0


@pytest.mark.parametrize("kind", ["pan_yaw", "stationary", "orbit"])
def test_camera_geometry_matches_reference(kind):
    from camc2v_b200 import camera, synth
    g = np.load(os.path.join(GOLD, "masks.npz"))
    K, w2c = synth.synth_camera(kind, T=16)
    torch.manual_seed(123)
    rel = camera.relative_c2w(w2c, torch.zeros(1, dtype=torch.long))
    Fm = camera.fundamental_matrices(K, rel)
    assert np.allclose(rel.numpy(), g[f"{kind}.rel_c2w"], rtol=1e-5, atol=1e-6)
    if np.array_equal(rel.numpy(), g[f"{kind}.rel_c2w"]):
        assert np.array_equal(Fm.numpy(), g[f"{kind}.F"]), "same poses must give the reference's F bit for bit (incl. the RNG draw)"


def test_flop_model_matches_reference_counts():
    from camc2v_b200.config import UNetConfig
    from camc2v_b200.flops import cfg_step_flops, unet_pass_flops
    cfg = UNetConfig()
    cond = unet_pass_flops(cfg, 1, 32, 845, False)
    unc = unet_pass_flops(cfg, 1, 32, 333, True)
    assert abs(cond["total"] / 1e12 - 7.875) < 0.01 and abs(unc["total"] / 1e12 - 7.121) < 0.01
    assert abs(cond["epipolar"] / 1e12 - 1.961) < 0.002 and abs(cond["ctx_kv"] / 1e12 - 0.691) < 0.002
    assert abs(cfg_step_flops(cfg, 1, 32) / 1e12 - 14.996) < 0.01
    assert abs(cfg_step_flops(cfg, 4, 32) / cfg_step_flops(cfg, 1, 32) - 4.0) < 1e-9


# ------------------------------------------------------------------------------------------------ multi-GPU host logic (gloo)
def _run_world(code: str, nproc: int, port: int):
    import tempfile
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
                            "127.0.0.1", "--master-port", str(port), path], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("ok") == nproc
    finally:
        os.remove(path)


def test_video_sharding_and_latent_gather_world2():
    """N > 1 path of bench.py / parallel.py: videos[rank::world] sharding, no step-time collective, one all_gather of the
    final latents — exercised with 2 gloo ranks on CPU."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from camc2v_b200 import parallel
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
vids = list(range(7))
mine = parallel.shard_videos(vids, rank, world)
assert mine == vids[rank::world]
lat = torch.stack([torch.full((4, 2, 2, 2), float(v)) for v in mine])
allv = parallel.gather_latents(lat, len(vids), rank, world)
if rank == 0:
    assert allv.shape == (7, 4, 2, 2, 2)
    assert [int(allv[i, 0, 0, 0, 0]) for i in range(7)] == vids
print("ok", rank)
''' % ROOT
    _run_world(code, 2, 29533)


def test_cfg_split_pair_exchange_world2():
    """CFG halves on a pair of ranks (parallel.CfgPair, BASELINE config 4): role 0 = cond, role 1 = uncond; after one
    2-rank all_gather both ranks hold (e_cond, e_uncond); both draw the same eta-noise; both see the pair's videos; the
    reference's update (oracle.ddim_oracle.cfg_ddim_update) then gives bit-identical latents on the two ranks."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from camc2v_b200 import parallel
from oracle import ddim_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
pair = parallel.make_cfg_pairs(rank, world)
assert pair.role == rank %% 2 and pair.pair_id == 0 and pair.n_pairs == 1
assert parallel.shard_videos_cfg_split(list(range(5)), rank, world) == list(range(5))
g = torch.Generator().manual_seed(7)
x = torch.randn(1, 4, 2, 4, 4, generator=g)
e_true = [torch.randn(1, 4, 2, 4, 4, generator=g) for _ in range(2)]       # what the cond / uncond pass would return
e_c, e_u = pair.exchange(e_true[pair.role].clone())
assert torch.equal(e_c, e_true[0]) and torch.equal(e_u, e_true[1])
n1, n2 = pair.noise(x.shape, torch.device("cpu")), pair.noise(x.shape, torch.device("cpu"))
assert not torch.equal(n1, n2)
s = ddim_oracle.ddim_schedule()
xp, _ = ddim_oracle.cfg_ddim_update(x, e_c, e_u, n1, float(s["alphas"][3]), float(s["alphas_prev"][3]), float(s["sigmas"][3]),
                                    float(s["sqrt_one_minus_alphas"][3]), 3.5, 0.7)
both = [torch.empty_like(xp) for _ in range(2)]
dist.all_gather(both, xp)
assert torch.equal(both[0], both[1]), "the two ranks of a pair must stay bit-identical"
print("ok", rank)
''' % ROOT
    _run_world(code, 2, 29534)


def test_emb_projection_stacking_matches_per_block_packs():
    """UNetModel stacks the emb_layers projections of all ResBlocks into one skinny GEMM (modules.EmbPack): every block's column
    slice of the stacked weight / bias must be exactly the block's own pack, and the slices must tile the matrix."""
    from camc2v_b200 import synth
    from camc2v_b200.modules import ResBlock, UNetConfig, build_unet
    unet = build_unet(UNetConfig(model_channels=64, origin_h=128, origin_w=128))
    synth.fill_module_(unet, seed=0)
    p = unet.pk()
    blocks = [m for m in unet.modules() if isinstance(m, ResBlock)]
    assert len(blocks) == 22
    off = 0
    for blk in blocks:
        o, c = blk._emb_slice
        assert o == off and c == blk.out_channels
        bp = blk.pk()
        assert torch.equal(p["we_all"][o:o + c], bp["we"]) and torch.equal(p["bemb_all"][o:o + c], bp["bemb"])
        off += c
    assert p["we_all"].shape == (off, 4 * 64) and p["bemb_all"].shape == (off,)


def test_integration_doc_maps_every_abi_symbol():
    """INTEGRATION.md names the reference interface each exported entry point replaces: no symbol of the header may be missing."""
    import re
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    hdr = open(os.path.join(root, "include", "camc2v_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(c2v_[a-z0-9_]+)\(", hdr)))
    assert len(syms) >= 35
    assert [s for s in syms if s not in doc] == []


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under camc2v_b200/ may import it (bench.py uses it only for the CPU baseline legs)."""
    import re
    pkg = os.path.join(os.path.abspath(os.path.join(os.path.dirname(__file__), "..")), "camc2v_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "oracle/" in src and f.endswith((".cu", ".cuh", ".h")) and "#include" in src and re.search(r'#include\s+"[^"]*oracle', src):
                    bad.append(f)
    assert bad == []


def test_unmodified_reference_constructor_accepts_the_cuda_unet():
    """INTEGRATION.md section 2 with NO other change: the reference's own CamContextI2V constructor, given
    `unet_config.target: camc2v_b200.modules.UNetModel`, runs its by-name forward re-binding and Epipolar injection
    (camcontexti2v.py:111-170) against this package's modules; the package declines the re-binding, adopts the injected modules
    as its own classes, and the reference UNet's state_dict loads with strict=True."""
    import os
    import sys
    import pytest
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, os.path.join(root, "oracle", "refgen"))
    import ref_harness as rh
    if not os.path.isdir(rh.REF_PKG):
        pytest.skip("reference sources not available")
    from camc2v_b200 import modules as M
    small = dict(model_channels=64)
    model = rh.build_reference_model(unet_overrides=small, unet_target="camc2v_b200.modules.UNetModel")
    unet = model.model.diffusion_model
    assert isinstance(unet, M.UNetModel)
    assert unet.__dict__.get("_declined_rebinds") == ["new_forward_for_unet"]
    assert type(unet).forward is M.UNetModel.forward and "forward" not in unet.__dict__
    blocks = [b for b in unet.modules() if isinstance(b, M.BasicTransformerBlock) and hasattr(b, "epipolar")]
    assert len(blocks) == 16 and all(isinstance(b.epipolar, M.Epipolar) and b.variant == "camcontext" for b in blocks)
    assert all("forward" not in b.__dict__ and "_forward" not in b.__dict__ for b in blocks)
    tts = [t for t in unet.modules() if isinstance(t, M.TemporalTransformer)]
    assert all(t.__dict__.get("_declined_rebinds") == ["new_forward_for_TemporalTransformer"] for t in tts)
    ref = rh.build_reference_model(unet_overrides=small)
    res = unet.load_state_dict(ref.model.diffusion_model.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the model object the samplers talk to is the reference's own LatentDiffusion
    assert type(model).__module__.startswith("model.") and hasattr(model, "apply_model") and model.parameterization == "eps"


def test_sampler_rejects_unsupported_reference_options():
    import pytest
    import torch
    from camc2v_b200.sampler import DDIMSampler

    class Stub:
        num_timesteps = 1000
        parameterization = "eps"
        use_dynamic_rescale = False
        alphas_cumprod = torch.linspace(0.999, 0.01, 1000)
        betas = torch.zeros(1000)

    s = DDIMSampler(Stub())
    s.make_schedule(25, "uniform_trailing", 1.0, verbose=False)
    x = torch.zeros(1, 4, 16, 8, 8)
    t = torch.zeros(1, dtype=torch.long)
    for kw in (dict(mask=torch.ones(1)), dict(x0=x), dict(paste_cond_frame=True), dict(noise_shaping=True), dict(timesteps=10),
               dict(precision=16), dict(quantize_denoised=True), dict(score_corrector=object())):
        with pytest.raises(NotImplementedError):
            s.p_sample_ddim(x, {}, t, index=0, **kw)
    with pytest.raises(NotImplementedError):
        s.sample(25, 1, (4, 16, 8, 8), mask=torch.ones(1))
    Stub.parameterization = "v"
    with pytest.raises(NotImplementedError):
        DDIMSampler(Stub())
