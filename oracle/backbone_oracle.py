"""CPU oracle for the network part of the Oryon hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as oryon_oracle.py: only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it).

A functional, float32, CPU-PyTorch restatement of ``Oryon.forward`` (reference net.py:142-167) driven by a
flat ``state_dict`` with the reference's own parameter names:

  vlm.clip_model.*          OpenAI CLIP ViT-L/14@336 towers.  THIRD-PARTY: ``clip==1.0`` (git install,
                            reference environment.yml:45) is not under /root/reference and not installed,
                            so its published architecture (clip/model.py: VisionTransformer, Transformer,
                            ResidualAttentionBlock, QuickGELU, LayerNorm-in-fp32) is restated here and
                            anchored on the reference's call sites models/vlm.py:43-61 and :63-86.  Pin:
                            tests/test_backbone_oracle.py checks it against the independent implementation in
                            ``transformers`` (CLIPVisionModel / CLIPTextModel, same weights).
  guidance_backbone.*       torchvision ``swin_b`` truncated at features.4 (reference net.py:45-75); the
                            library model itself is run (torchvision is a dependency of the reference too).
  fusion.*                  reference models/fusion.py:577-625 (+ blocks :40-235, :240-266, :301-332, :386-434)
  decoder.*                 reference models/decoder.py:82-108

Pin for fusion/decoder: tests/golden/backbone_*.npz were produced by the UNMODIFIED reference modules
(oracle/make_golden_backbone.py, run in the build container) from the same seeded state_dict; this file
must reproduce them (tests/test_backbone_oracle.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _ln(x: Tensor, w: Dict[str, Tensor], name: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], eps)


def _lin(x: Tensor, w: Dict[str, Tensor], name: str) -> Tensor:
    return F.linear(x, w[name + ".weight"], w.get(name + ".bias"))


# --------------------------------------------------------------------------------------------
# CLIP (clip/model.py restated; call sites models/vlm.py:43-61, :63-86)
# --------------------------------------------------------------------------------------------


def clip_resblock(x: Tensor, w: Dict[str, Tensor], p: str, heads: int, causal: bool) -> Tensor:
    """``x + attn(ln_1(x)); x + mlp(ln_2(x))`` with ``nn.MultiheadAttention`` semantics (packed in_proj,
    q scaled by d^-0.5, additive -inf upper-triangular mask for the text tower) and the QuickGELU MLP.
    ``x`` is ``[N, L, D]`` (batch first; the reference permutes to LND only because nn.MultiheadAttention
    wants it, vlm.py:53-55)."""
    n, l, d = x.shape
    h = _ln(x, w, p + ".ln_1")
    qkv = F.linear(h, w[p + ".attn.in_proj_weight"], w[p + ".attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    hd = d // heads
    q = q.view(n, l, heads, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.view(n, l, heads, hd).transpose(1, 2)
    v = v.view(n, l, heads, hd).transpose(1, 2)
    att = q @ k.transpose(-1, -2)
    if causal:
        att = att + torch.full((l, l), float("-inf")).triu_(1)
    att = torch.softmax(att, dim=-1)
    o = (att @ v).transpose(1, 2).reshape(n, l, d)
    x = x + _lin(o, w, p + ".attn.out_proj")
    h = _ln(x, w, p + ".ln_2")
    h = _lin(h, w, p + ".mlp.c_fc")
    h = h * torch.sigmoid(1.702 * h)  # QuickGELU
    return x + _lin(h, w, p + ".mlp.c_proj")


def clip_preprocess(image: Tensor, size: int = 336) -> Tensor:
    """``Compose([Resize(size, BICUBIC), CenterCrop(size), Normalize])`` on a float tensor (vlm.py:21-22, :45);
    torchvision 0.13 tensors: bicubic a=-0.75, align_corners=False, no antialias."""
    x = F.interpolate(image, size=(size, size), mode="bicubic", align_corners=False, antialias=False)
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def clip_encode_image(w: Dict[str, Tensor], image: Tensor, *, layers: int = 24, heads: int = 16, patch: int = 14,
                      size: int = 336) -> Tensor:
    """``CLIPEncoder.encode_image`` (vlm.py:43-61): ``[B,3,224,224]`` in [0,1] -> ``[B,width,grid,grid]``
    (``ln_post`` of the patch tokens, no ``visual.proj``)."""
    p = "vlm.clip_model.visual"
    x = clip_preprocess(image, size)
    x = F.conv2d(x, w[p + ".conv1.weight"], None, stride=patch)
    b, width, g, _ = x.shape
    x = x.reshape(b, width, -1).permute(0, 2, 1)
    cls = w[p + ".class_embedding"] + torch.zeros(b, 1, width)
    x = torch.cat([cls, x], dim=1) + w[p + ".positional_embedding"]
    x = _ln(x, w, p + ".ln_pre")
    for i in range(layers):
        x = clip_resblock(x, w, f"{p}.transformer.resblocks.{i}", heads, causal=False)
    toks = _ln(x[:, 1:, :], w, p + ".ln_post")
    return toks.transpose(1, 2).reshape(b, width, g, g)


def clip_encode_tokens(w: Dict[str, Tensor], tokens: Tensor, *, layers: int = 12, heads: int = 12) -> Tensor:
    """Text tower on already-tokenised prompts ``[n,77]`` int64 -> ``[n,embed]`` (vlm.py:74-83): token +
    positional embedding, causal transformer, ``ln_final``, feature at the EOT position
    (``argmax`` of the token ids), ``@ text_projection``."""
    p = "vlm.clip_model"
    x = w[p + ".token_embedding.weight"][tokens] + w[p + ".positional_embedding"]
    for i in range(layers):
        x = clip_resblock(x, w, f"{p}.transformer.resblocks.{i}", heads, causal=True)
    x = _ln(x, w, p + ".ln_final")
    eot = tokens.argmax(dim=-1)
    x = x[torch.arange(x.shape[0]), eot]
    return x @ w[p + ".text_projection"]


# --------------------------------------------------------------------------------------------
# guidance backbone (net.py:45-75)
# --------------------------------------------------------------------------------------------


def guidance_backbone(w: Dict[str, Tensor]):
    """torchvision ``swin_b`` truncated by ``create_feature_extractor`` at the reference's return nodes,
    loaded with the ``guidance_backbone.*`` entries of ``w`` (net.py:45-58)."""
    from torchvision.models import swin_b
    from torchvision.models.feature_extraction import create_feature_extractor
    swin = swin_b(weights=None)
    nodes = {"features.1.1.add_1": "guidance3", "features.2.reduction": "guidance2", "features.4.reduction": "guidance1"}
    net = create_feature_extractor(swin, return_nodes=nodes)
    sd = {k[len("guidance_backbone."):]: v for k, v in w.items() if k.startswith("guidance_backbone.")}
    own = net.state_dict()
    missing = [k for k in own if k not in sd]
    assert not missing, missing[:5]
    net.load_state_dict({k: sd[k] for k in own}, strict=True)
    return net.eval()


def guidance_embeds(net, image: Tensor) -> List[Tensor]:
    """``Oryon.get_guidance_embeds`` (net.py:60-75): bicubic 384 ``align_corners=True``, ImageNet
    normalisation, truncated Swin, NHWC -> NCHW.  Returns ``[guid1 [B,512,24,24], guid2 [B,256,48,48],
    guid3 [B,128,96,96]]``."""
    x = F.interpolate(image.clone(), size=(384, 384), mode="bicubic", align_corners=True)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    x = (x - mean) / std
    with torch.no_grad():
        outs = net(x)
    return [outs[k].permute(0, 3, 1, 2).contiguous() for k in ("guidance1", "guidance2", "guidance3")]


# --------------------------------------------------------------------------------------------
# fusion (models/fusion.py)
# --------------------------------------------------------------------------------------------


def _window_partition(x: Tensor, ws: int) -> Tensor:
    b, h, wd, c = x.shape
    x = x.view(b, h // ws, ws, wd // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def _window_reverse(win: Tensor, ws: int, h: int, wd: int) -> Tensor:
    b = int(win.shape[0] / (h * wd / ws / ws))
    x = win.view(b, h // ws, wd // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(b, h, wd, -1)


def _fusion_shift_mask(h: int, wd: int, ws: int, shift: int) -> Tensor:
    """fusion.py:147-167: region ids on the un-padded grid, -100 (not -inf) between regions."""
    img = torch.zeros((1, h, wd, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = _window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


def fusion_swin_block(x: Tensor, guid: Tensor, w: Dict[str, Tensor], p: str, *, res: Tuple[int, int], ws: int, shift: int,
                      heads: int) -> Tensor:
    """``SwinTransformerBlock.forward`` + ``WindowAttention.forward`` (fusion.py:169-213, :76-103): q and k
    from ``cat(norm1(x), guidance)`` (256-d), v from the first ``dim`` channels only, no relative-position bias."""
    h, wd = res
    b, l, c = x.shape
    shortcut = x
    y = _ln(x, w, p + ".norm1").view(b, h, wd, c)
    y = torch.cat([y, guid.view(b, h, wd, -1)], dim=-1)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    xw = _window_partition(y, ws).view(-1, ws * ws, y.shape[-1])
    b_, n, _ = xw.shape
    q = _lin(xw, w, p + ".attn.q").reshape(b_, n, heads, -1).permute(0, 2, 1, 3)
    k = _lin(xw, w, p + ".attn.k").reshape(b_, n, heads, -1).permute(0, 2, 1, 3)
    v = _lin(xw[:, :, :c], w, p + ".attn.v").reshape(b_, n, heads, -1).permute(0, 2, 1, 3)
    q = q * ((c // heads) ** -0.5)
    attn = q @ k.transpose(-2, -1)
    if shift > 0:
        mask = _fusion_shift_mask(h, wd, ws, shift)
        nw = mask.shape[0]
        attn = attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(b_, n, -1)
    o = _lin(o, w, p + ".attn.proj").view(-1, ws, ws, c)
    y = _window_reverse(o, ws, h, wd)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = shortcut + y.view(b, h * wd, c)
    m = _lin(F.gelu(_lin(_ln(x, w, p + ".norm2"), w, p + ".mlp.fc1")), w, p + ".mlp.fc2")  # timm Mlp: fc1 -> GELU(erf) -> fc2
    return x + m


def fusion_class_transformer(x: Tensor, text_guid: Tensor, w: Dict[str, Tensor], p: str, *, pool: int, heads: int) -> Tensor:
    """``ClassTransformerLayer.forward`` (fusion.py:409-434) with ``AttentionLayer`` (:318-332) and
    ``LinearAttention`` (:246-266).  ``x [B,C,T,H,W]``, ``text_guid [B,T,C]``."""
    b, c, t, h, wd = x.shape
    xp = F.avg_pool2d(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, wd), pool)
    hp, wp = xp.shape[-2:]
    xp = xp.view(b, t, c, hp, wp).permute(0, 3, 4, 1, 2).reshape(b * hp * wp, t, c)           # (B H W) T C
    g = text_guid[:, None, None].expand(b, hp, wp, t, text_guid.shape[-1]).reshape(b * hp * wp, t, -1)
    y = _ln(xp, w, p + ".norm1")
    q = _lin(torch.cat([y, g], dim=-1), w, p + ".attention.q").view(-1, t, heads, c // heads)
    k = _lin(torch.cat([y, g], dim=-1), w, p + ".attention.k").view(-1, t, heads, c // heads)
    v = _lin(y, w, p + ".attention.v").view(-1, t, heads, c // heads)
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    vl = v.size(1)
    v = v / vl
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + 1e-6)
    att = (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * vl).reshape(-1, t, c)
    xp = xp + att
    xp = xp + _lin(torch.relu(_lin(_ln(xp, w, p + ".norm2"), w, p + ".MLP.0")), w, p + ".MLP.2")
    xp = xp.view(b, hp, wp, t, c).permute(0, 3, 4, 1, 2).reshape(b * t, c, hp, wp)
    xp = F.interpolate(xp, size=(h, wd), mode="bilinear", align_corners=True)
    xp = xp.view(b, t, c, h, wd).permute(0, 2, 1, 3, 4)
    return x + xp


def fusion_forward(w: Dict[str, Tensor], img_feats: Tensor, text_feats: Tensor, guid1: Tensor, *, layers: int = 2,
                   heads: int = 4, ws: int = 12, pool: int = 6) -> Tensor:
    """``ImageTextFusion.forward`` (fusion.py:602-625).  ``img_feats [B,1024,24,24]``, ``text_feats [B,1,80,768]``,
    ``guid1 [B,512,24,24]`` -> ``[B,128,1,24,24]``."""
    p = "fusion"
    b, d, h, wd = img_feats.shape
    x = F.conv1d(img_feats.reshape(b, d, h * wd), w[p + ".clip_conv.weight"], w[p + ".clip_conv.bias"]).reshape(b, -1, h, wd)
    xi = F.normalize(x, dim=1)
    tf = F.normalize(text_feats, dim=-1)
    corr = torch.einsum("bchw,btpc->bpthw", xi, tf)                                            # B P T H W
    t = corr.shape[2]
    ce = F.conv2d(corr.permute(0, 2, 1, 3, 4).reshape(b * t, -1, h, wd), w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=3)
    ce = ce.view(b, t, -1, h, wd).permute(0, 2, 1, 3, 4)                                        # B C T H W
    pg = torch.relu(F.conv2d(guid1, w[p + ".guidance_projection.0.weight"], w[p + ".guidance_projection.0.bias"], padding=1))
    tg = text_feats.mean(dim=-2)
    tg = tg / tg.norm(dim=-1, keepdim=True)
    tg = torch.relu(_lin(tg, w, p + ".text_guidance_projection.0"))                             # B T 128
    c = ce.shape[1]
    for li in range(layers):
        lp = f"{p}.layers.{li}"
        xs = ce.permute(0, 2, 3, 4, 1).reshape(b * t, h * wd, c)                                # (B T) (H W) C
        g = pg[:, None].expand(b, t, pg.shape[1], h, wd).permute(0, 1, 3, 4, 2).reshape(b * t, h * wd, -1)
        g = _ln(g, w, lp + ".swin_block.guidance_norm")
        xs = fusion_swin_block(xs, g, w, lp + ".swin_block.block_1", res=(h, wd), ws=ws, shift=0, heads=heads)
        xs = fusion_swin_block(xs, g, w, lp + ".swin_block.block_2", res=(h, wd), ws=ws, shift=ws // 2, heads=heads)
        ce = xs.view(b, t, h, wd, c).permute(0, 4, 1, 2, 3)
        ce = fusion_class_transformer(ce, tg, w, lp + ".attention", pool=pool, heads=heads)
    return ce


# --------------------------------------------------------------------------------------------
# decoder (models/decoder.py)
# --------------------------------------------------------------------------------------------


def _double_conv(x: Tensor, w: Dict[str, Tensor], p: str) -> Tensor:
    """``DoubleConv`` (decoder.py:9-26): (conv3x3 no bias -> GroupNorm(C/16 groups) -> ReLU) x 2."""
    for ci, gi in ((0, 1), (3, 4)):
        x = F.conv2d(x, w[f"{p}.double_conv.{ci}.weight"], None, padding=1)
        gw = w[f"{p}.double_conv.{gi}.weight"]
        x = torch.relu(F.group_norm(x, gw.shape[0] // 16, gw, w[f"{p}.double_conv.{gi}.bias"], 1e-5))
    return x


def _up(x: Tensor, guid, w: Dict[str, Tensor], p: str) -> Tensor:
    """``Up`` (decoder.py:29-42): ConvTranspose2d(k=2, s=2) -> cat guidance -> DoubleConv."""
    x = F.conv_transpose2d(x, w[p + ".up.weight"], w[p + ".up.bias"], stride=2)
    if guid is not None:
        x = torch.cat([x, guid], dim=1)
    return _double_conv(x, w, p + ".conv")


def decoder_forward(w: Dict[str, Tensor], x: Tensor, guidance: List[Tensor]) -> Tuple[Tensor, Tensor]:
    """``StandardDecoder.forward`` (decoder.py:82-108), ``extra_upsampling=True``, ``use_guidance=True``:
    ``x [B,128,1,24,24]``, ``guidance = [guid1, guid2, guid3]`` -> ``(mask logits [B,1,192,192], featmap [B,32,192,192])``."""
    p = "decoder"
    pg = [torch.relu(F.conv2d(g, w[f"{p}.decoder_guidance_projection.{i}.0.weight"], w[f"{p}.decoder_guidance_projection.{i}.0.bias"],
                              padding=1)) for i, g in enumerate(guidance[1:])]
    b, c, t, h, wd = x.shape
    ce = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, wd)
    ce = _up(ce, pg[0], w, p + ".decoder1")
    ce = _up(ce, pg[1], w, p + ".decoder2")
    ce = _up(ce, None, w, p + ".decoder3")
    featmap = ce.view(b, t * ce.shape[1], *ce.shape[-2:]).clone()
    logits = F.conv2d(ce, w[p + ".head.weight"], w[p + ".head.bias"], padding=1).view(b, t, *ce.shape[-2:])
    return logits, featmap


# --------------------------------------------------------------------------------------------
# the whole network (net.py:142-167)
# --------------------------------------------------------------------------------------------


def oryon_forward(w: Dict[str, Tensor], rgb_a: Tensor, rgb_q: Tensor, tokens: Tensor, *, swin=None, vis_layers: int = 24,
                  txt_layers: int = 12, return_stages: bool = False):
    """``Oryon.forward`` on tensors: ``rgb_a/q [B,3,224,224]`` in [0,1], ``tokens [B,80,77]`` (the tokenised
    80 templated prompts, i.e. after ``prompt_list[1:]`` of vlm.py:67) -> the reference's output dict."""
    if swin is None:
        swin = guidance_backbone(w)
    b, tcount = tokens.shape[:2]
    with torch.no_grad():
        vis_a = clip_encode_image(w, rgb_a, layers=vis_layers)
        vis_q = clip_encode_image(w, rgb_q, layers=vis_layers)
        prompt = clip_encode_tokens(w, tokens.reshape(b * tcount, -1), layers=txt_layers).view(b, tcount, -1).unsqueeze(1)
        ga, gq = guidance_embeds(swin, rgb_a), guidance_embeds(swin, rgb_q)
        fa = fusion_forward(w, vis_a, prompt, ga[0])
        fq = fusion_forward(w, vis_q, prompt, gq[0])
        mask_a, feat_a = decoder_forward(w, fa, ga)
        mask_q, feat_q = decoder_forward(w, fq, gq)
    out = dict(featmap_a=feat_a, featmap_q=feat_q, mask_a=mask_a, mask_q=mask_q)
    if return_stages:
        out.update(clip_a=vis_a, clip_q=vis_q, prompt=prompt, guid_a=ga, guid_q=gq, fusion_a=fa, fusion_q=fq)
    return out
