"""Seeded random ``state_dict`` for the whole network with the reference's parameter names and shapes
(``Oryon``: net.py:24-36; CLIP ViT-L/14@336 as ``clip.load`` builds it, vlm.py:19; torchvision ``swin_b``
truncated at ``features.4``, net.py:45-58; ``ImageTextFusion`` fusion.py:518-574; ``StandardDecoder``
decoder.py:49-80).  No checkpoint is available offline, so parity and benchmarks run on
these weights; the scales are chosen so that activations stay O(1) through all 24 + 12 transformer layers and
attention is not uniform (every code path is numerically exercised).  Used by tests, bench.py and
oracle/make_golden_backbone.py (which checks that the reference modules accept it with ``strict=True``).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
from torch import Tensor

CLIP_VIS = dict(width=1024, layers=24, heads=16, patch=14, grid=24)
CLIP_TXT = dict(width=768, layers=12, heads=12, ctx=77, vocab=49408, embed=768)


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


class _Maker:
    def __init__(self, seed: int):
        self.g = _gen(seed)
        self.sd: Dict[str, Tensor] = {}

    def normal(self, name, shape, std):
        self.sd[name] = torch.randn(*shape, generator=self.g) * std

    def linear(self, name, cout, cin, std=None, bias=True, bias_std=0.02):
        self.normal(name + ".weight", (cout, cin), std if std is not None else cin ** -0.5)
        if bias:
            self.normal(name + ".bias", (cout,), bias_std)

    def conv(self, name, cout, cin, k, bias=True, std=None):
        self.normal(name + ".weight", (cout, cin, k, k), std if std is not None else (cin * k * k) ** -0.5)
        if bias:
            self.normal(name + ".bias", (cout,), 0.02)

    def norm(self, name, c):
        self.sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=self.g)
        self.sd[name + ".bias"] = 0.05 * torch.randn(c, generator=self.g)


def _clip_blocks(m: _Maker, prefix: str, width: int, layers: int):
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        m.norm(p + ".ln_1", width)
        m.normal(p + ".attn.in_proj_weight", (3 * width, width), width ** -0.5)
        m.normal(p + ".attn.in_proj_bias", (3 * width,), 0.02)
        m.linear(p + ".attn.out_proj", width, width, std=proj_std)
        m.norm(p + ".ln_2", width)
        m.linear(p + ".mlp.c_fc", 4 * width, width, std=(2 * width) ** -0.5)
        m.linear(p + ".mlp.c_proj", width, 4 * width, std=proj_std)


def clip_state_dict(seed: int, vis_layers: int = 24, txt_layers: int = 12) -> Dict[str, Tensor]:
    m = _Maker(seed)
    v, t = CLIP_VIS, CLIP_TXT
    p = "vlm.clip_model.visual"
    m.normal(p + ".conv1.weight", (v["width"], 3, v["patch"], v["patch"]), (3 * v["patch"] ** 2) ** -0.5)
    m.normal(p + ".class_embedding", (v["width"],), v["width"] ** -0.5)
    m.normal(p + ".positional_embedding", (v["grid"] ** 2 + 1, v["width"]), v["width"] ** -0.5)
    m.norm(p + ".ln_pre", v["width"])
    _clip_blocks(m, p + ".transformer", v["width"], vis_layers)
    m.norm(p + ".ln_post", v["width"])
    p = "vlm.clip_model"
    m.normal(p + ".token_embedding.weight", (t["vocab"], t["width"]), 0.02)
    m.normal(p + ".positional_embedding", (t["ctx"], t["width"]), 0.01)
    _clip_blocks(m, p + ".transformer", t["width"], txt_layers)
    m.norm(p + ".ln_final", t["width"])
    m.normal(p + ".text_projection", (t["width"], t["embed"]), t["width"] ** -0.5)
    return m.sd


def _rel_pos_index(ws: int) -> Tensor:
    """torchvision ShiftedWindowAttention.define_relative_position_index."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1).flatten()


def swin_state_dict(seed: int) -> Dict[str, Tensor]:
    """The part of torchvision ``swin_b`` that survives ``create_feature_extractor`` (features.0 .. features.4)."""
    m = _Maker(seed)
    p = "guidance_backbone.features"
    m.conv(p + ".0.0", 128, 3, 4)
    m.norm(p + ".0.2", 128)
    for stage, dim, heads in ((1, 128, 4), (3, 256, 8)):
        for blk in range(2):
            b = f"{p}.{stage}.{blk}"
            m.norm(b + ".norm1", dim)
            m.normal(b + ".attn.relative_position_bias_table", (169, heads), 0.5)
            m.sd[b + ".attn.relative_position_index"] = _rel_pos_index(7)
            m.linear(b + ".attn.qkv", 3 * dim, dim)
            m.linear(b + ".attn.proj", dim, dim, std=0.5 * dim ** -0.5)
            m.norm(b + ".norm2", dim)
            m.linear(b + ".mlp.0", 4 * dim, dim)
            m.linear(b + ".mlp.3", dim, 4 * dim, std=0.5 * (4 * dim) ** -0.5)
        m.linear(f"{p}.{stage + 1}.reduction", 2 * dim, 4 * dim, bias=False)
        m.norm(f"{p}.{stage + 1}.norm", 4 * dim)
    return m.sd


def fusion_decoder_state_dict(seed: int) -> Dict[str, Tensor]:
    m = _Maker(seed)
    for li in range(2):
        lp = f"fusion.layers.{li}"
        for blk in ("block_1", "block_2"):
            b = f"{lp}.swin_block.{blk}"
            m.norm(b + ".norm1", 128)
            m.linear(b + ".attn.q", 128, 256)
            m.linear(b + ".attn.k", 128, 256)
            m.linear(b + ".attn.v", 128, 128)
            m.linear(b + ".attn.proj", 128, 128, std=0.5 * 128 ** -0.5)
            m.norm(b + ".norm2", 128)
            m.linear(b + ".mlp.fc1", 512, 128)
            m.linear(b + ".mlp.fc2", 128, 512, std=0.5 * 512 ** -0.5)
        m.norm(lp + ".swin_block.guidance_norm", 128)
        a = lp + ".attention"
        m.linear(a + ".attention.q", 128, 256)
        m.linear(a + ".attention.k", 128, 256)
        m.linear(a + ".attention.v", 128, 128)
        m.linear(a + ".MLP.0", 512, 128)
        m.linear(a + ".MLP.2", 128, 512, std=0.5 * 512 ** -0.5)
        m.norm(a + ".norm1", 128)
        m.norm(a + ".norm2", 128)
    # the cosine cost volume is O(1/sqrt(768)) for random features: a larger conv1 keeps its output O(1)
    m.conv("fusion.conv1", 128, 80, 7, std=28.0 * (80 * 49) ** -0.5)
    m.normal("fusion.clip_conv.weight", (768, 1024, 1), 1024 ** -0.5)
    m.normal("fusion.clip_conv.bias", (768,), 0.02)
    m.conv("fusion.guidance_projection.0", 128, 512, 3)
    m.linear("fusion.text_guidance_projection.0", 128, 768, std=6.0 * 768 ** -0.5)
    m.conv("decoder.decoder_guidance_projection.0.0", 32, 256, 3)
    m.conv("decoder.decoder_guidance_projection.1.0", 16, 128, 3)
    for name, cin, cup, cmid in (("decoder1", 128, 96, 64), ("decoder2", 64, 48, 32), ("decoder3", 32, 32, 32)):
        d = f"decoder.{name}"
        m.normal(d + ".up.weight", (cin, cup, 2, 2), cin ** -0.5)
        m.normal(d + ".up.bias", (cup,), 0.02)
        m.conv(d + ".conv.double_conv.0", cmid, cin if name != "decoder3" else 32, 3, bias=False)
        m.norm(d + ".conv.double_conv.1", cmid)
        m.conv(d + ".conv.double_conv.3", cmid, cmid, 3, bias=False)
        m.norm(d + ".conv.double_conv.4", cmid)
    m.conv("decoder.head", 1, 32, 3)
    return m.sd


def oryon_state_dict(seed: int, vis_layers: int = 24, txt_layers: int = 12) -> Dict[str, Tensor]:
    sd = clip_state_dict(seed, vis_layers, txt_layers)
    sd.update(swin_state_dict(seed + 1))
    sd.update(fusion_decoder_state_dict(seed + 2))
    return sd


def synthetic_tokens(seed: int, b: int, n_prompts: int = 80, ctx: int = 77, vocab: int = 49408) -> Tensor:
    """Token ids shaped like ``SimpleTokenizer`` output (tokenizer.py:136-151): SOT (49406), 4-20 word tokens,
    EOT (49407, the arg-max id, vlm.py:81), zero padding.  ``[b, n_prompts, ctx]`` int64."""
    g = _gen(seed)
    out = torch.zeros(b, n_prompts, ctx, dtype=torch.int64)
    for i in range(b):
        for j in range(n_prompts):
            n = int(torch.randint(4, 21, (1,), generator=g))
            out[i, j, 0] = vocab - 2
            out[i, j, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=g)
            out[i, j, 1 + n] = vocab - 1
    return out


def synthetic_images(seed: int, b: int, size: int = 224) -> Tensor:
    """Smooth-ish RGB in [0,1]: low-frequency pattern + noise, ``[b,3,size,size]``."""
    g = _gen(seed)
    coarse = torch.rand(b, 3, 14, 14, generator=g)
    x = torch.nn.functional.interpolate(coarse, size=(size, size), mode="bilinear", align_corners=False)
    return (0.8 * x + 0.2 * torch.rand(b, 3, size, size, generator=g)).clamp(0, 1)
