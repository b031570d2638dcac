"""Hyper-parameters and block topology of the lvdm 3D-UNet on the DDIM/CFG hot path.

Mirrors the constructor arguments of the reference `UNetModel`
(R/lvdm/modules/networks/openaimodel3d.py:311-342) as fixed by
configs/models/camcontexti2v_256.yaml:40-69, and re-derives the block list that the constructor
builds (openaimodel3d.py:384-565) as plain data, so that the CUDA modules and the CPU oracle walk the
same topology and produce the same `state_dict` keys.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    model_channels: int = 320
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_head_channels: int = 64
    context_dim: int = 1024
    temporal_length: int = 16
    init_attn_heads: int = 8          # openaimodel3d.py:389-402 (n_heads=8 hard-coded)
    default_fs: int = 3
    text_context_len: int = 77        # attention.py:49
    # camera branch (configs/models/camcontexti2v_256.yaml:152-160, camcontexti2v.py:151-168)
    epipolar: bool = True
    pluker_projection: bool = True
    num_register_tokens: int = 4
    origin_h: int = 256
    origin_w: int = 256
    # variant block: "camcontext" (== CamI2V), "cameractrl", "motionctrl", "none"
    variant: str = "camcontext"

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.model_channels


@dataclass
class Layer:
    kind: str                 # conv_in | res | spatial | temporal | down | up
    name: str                 # state_dict prefix, e.g. "input_blocks.1.0"
    cin: int = 0
    cout: int = 0
    heads: int = 0
    ds: int = 1
    epipolar: bool = False    # temporal block carries Epipolar + pluker_projection
    conv_proj: bool = False   # init_attn: proj_in/out are Conv1d(k=1) (openaimodel3d.py:391 -> use_linear=False)


@dataclass
class Block:
    name: str
    layers: List[Layer] = field(default_factory=list)
    ds: int = 1               # ds value recorded in input_ds/output_ds for this block


@dataclass
class Topology:
    input_blocks: List[Block]
    init_attn: Layer
    middle: Block
    output_blocks: List[Block]
    out_channels_last: int


def build_topology(cfg: UNetConfig) -> Topology:
    mc = cfg.model_channels
    hc = cfg.num_head_channels
    ep = cfg.epipolar

    init_inner = cfg.init_attn_heads * hc

    def attn_layers(prefix: str, start: int, ch: int, ds: int) -> List[Layer]:
        # camcontexti2v.py:140-142: a temporal block gets Epipolar/pluker_projection only if its
        # width differs from init_attn's inner width (8*64 = 512).
        return [
            Layer("spatial", f"{prefix}.{start}", ch, ch, ch // hc, ds),
            Layer("temporal", f"{prefix}.{start + 1}", ch, ch, ch // hc, ds, epipolar=ep and ch != init_inner),
        ]

    inputs: List[Block] = [Block("input_blocks.0", [Layer("conv_in", "input_blocks.0.0", cfg.in_channels, mc)], 1)]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            name = f"input_blocks.{len(inputs)}"
            layers = [Layer("res", f"{name}.0", ch, mult * mc, ds=ds)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers += attn_layers(name, 1, ch, ds)
            inputs.append(Block(name, layers, ds))
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            name = f"input_blocks.{len(inputs)}"
            inputs.append(Block(name, [Layer("down", f"{name}.0", ch, ch, ds=ds)], ds))
            chans.append(ch)
            ds *= 2

    init_attn = Layer("temporal", "init_attn.0", mc, mc, cfg.init_attn_heads, 1, epipolar=False, conv_proj=True)

    middle = Block("middle_block", [Layer("res", "middle_block.0", ch, ch, ds=ds)]
                   + attn_layers("middle_block", 1, ch, ds)
                   + [Layer("res", "middle_block.3", ch, ch, ds=ds)], ds)

    outputs: List[Block] = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            name = f"output_blocks.{len(outputs)}"
            layers = [Layer("res", f"{name}.0", ch + ich, mult * mc, ds=ds)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers += attn_layers(name, 1, ch, ds)
            blk_ds = ds
            if level and i == cfg.num_res_blocks:
                layers.append(Layer("up", f"{name}.{len(layers)}", ch, ch, ds=ds))
                ds //= 2
            outputs.append(Block(name, layers, blk_ds))
    return Topology(inputs, init_attn, middle, outputs, ch)
