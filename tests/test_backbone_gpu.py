"""GPU parity of the network (oryon_backbone_forward / oryon_text_forward) against the CPU oracle
(oracle/backbone_oracle.py, pinned to the reference modules) at the reference's full size: CLIP ViT-L/14@336
(24 layers), text tower (12 layers), truncated swin_b, fusion, decoder; seeded random weights and inputs.

Tolerance: BASELINE.json north_star -- "outputs match the reference within 1e-3 on feature maps" -- absolute, on
every element of every stage (all stages are O(1) by construction of the weights, so this is also ~1e-3 relative).
Both parity modes must meet it: the default (precision 2: fp16 split pairs with three tcgen05 products per product, except the four
linear layers of every CLIP vision block, whose cross terms run as one 8-bit product) and precision 3 (three products everywhere);
precision 1 is checked to 5e-2.
"""
import numpy as np
import pytest
import torch

import backbone_oracle as bo
from gpu_util import need_gpu
from oryon_b200 import synth_backbone as sb
from oryon_b200.net import Oryon

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def setup():
    need_gpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    w = sb.oryon_state_dict(11)
    rgb_a, rgb_q = sb.synthetic_images(1, 2), sb.synthetic_images(2, 2)
    tokens = sb.synthetic_tokens(3, 2)
    tokens[1] = tokens[0]  # second pair reuses the first prompt set in the oracle (halves its CPU time)
    with torch.no_grad():
        ref = bo.oryon_forward(w, rgb_a, rgb_q, tokens[:1].expand(2, -1, -1), return_stages=True)
    model = Oryon(None, "cuda:0", state_dict=w)
    return w, rgb_a, rgb_q, tokens, ref, model


def _maxerr(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def test_text_tower(setup):
    w, _, _, tokens, ref, model = setup
    emb = model.encode_tokens(tokens[0])
    err = _maxerr(emb, ref["prompt"][0, 0])
    print(f"text tower max abs err {err:.3e}")
    assert err < TOL


def test_network_stages_and_outputs(setup):
    w, rgb_a, rgb_q, tokens, ref, model = setup
    emb = model.encode_tokens(tokens[0])[None].expand(2, -1, -1).contiguous()
    out, dbg = model.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb, return_debug=True)
    torch.cuda.synchronize()
    B = 2
    errs = {}
    errs["clip"] = max(_maxerr(dbg["clip_tokens"][:B], ref["clip_a"]), _maxerr(dbg["clip_tokens"][B:], ref["clip_q"]))
    for i, name in enumerate(("guid1", "guid2", "guid3")):
        errs[name] = max(_maxerr(dbg[name][:B], ref["guid_a"][i]), _maxerr(dbg[name][B:], ref["guid_q"][i]))
    errs["fusion"] = max(_maxerr(dbg["fusion"][:B], ref["fusion_a"][:, :, 0]), _maxerr(dbg["fusion"][B:], ref["fusion_q"][:, :, 0]))
    for k in ("featmap_a", "featmap_q", "mask_a", "mask_q"):
        errs[k] = _maxerr(out[k], ref[k])
    print("max abs errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert out["featmap_a"].shape == (2, 32, 192, 192) and out["mask_a"].shape == (2, 1, 192, 192)
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    # predicted masks (losses.py:56-59: sigmoid > 0.5 == logit > 0) agree except where the logit is within the tolerance of 0
    for k in ("mask_a", "mask_q"):
        mine, theirs = out[k].cpu() > 0, ref[k] > 0
        differ = mine != theirs
        assert (ref[k][differ].abs() < TOL).all()


def test_forward_dict_interface_and_batch_invariance(setup):
    """Reference call shape (net.py:142): a batch dict in, the four maps out; the result for a pair does not depend
    on what else is in the batch nor on the chunking (max_pairs_per_pass)."""
    w, rgb_a, rgb_q, tokens, ref, model = setup
    xs = {"anchor": {"rgb": rgb_a.cuda()}, "query": {"rgb": rgb_q.cuda()}, "prompt_tokens": tokens.cuda()}
    out = model(xs)
    single = model({"anchor": {"rgb": rgb_a[1:].cuda()}, "query": {"rgb": rgb_q[1:].cuda()}, "prompt_tokens": tokens[1:].cuda()})
    for k in ("featmap_a", "featmap_q", "mask_a", "mask_q"):
        assert _maxerr(out[k][1:], single[k]) < 1e-5, k
        assert _maxerr(out[k], ref[k]) < TOL


def test_attention_kernels_agree(setup, monkeypatch):
    """The default fused attention (two softmax groups on alternating key tiles, partial results merged at the end) against the
    one-group online kernel (ORYON_ATTN_LOCKSTEP=1, read per call): different summation order and reference maxima, so not
    bit-identical; the CLIP tokens and the network outputs agree to 1e-4 (measured 2-5e-5), inside the oracle gate both are held to."""
    w, rgb_a, rgb_q, tokens, ref, model = setup
    emb = model.encode_tokens(tokens[0])[None].expand(2, -1, -1).contiguous()
    out, dbg = model.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb, return_debug=True)
    torch.cuda.synchronize()
    monkeypatch.setenv("ORYON_ATTN_LOCKSTEP", "1")
    out1, dbg1 = model.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb, return_debug=True)
    torch.cuda.synchronize()
    gaps = {k: _maxerr(out[k], out1[k]) for k in ("featmap_a", "featmap_q", "mask_a", "mask_q")}
    gaps["clip_tokens"] = _maxerr(dbg["clip_tokens"], dbg1["clip_tokens"])
    print("ping-pong vs lockstep:", {k: f"{v:.2e}" for k, v in gaps.items()})
    assert all(torch.isfinite(v).all() for v in out.values())
    assert all(v < 1e-4 for v in gaps.values()), gaps
    errs1 = {k: _maxerr(out1[k], ref[k]) for k in ("featmap_a", "featmap_q", "mask_a", "mask_q")}
    assert all(v < TOL for v in errs1.values()), errs1


def test_three_products_everywhere_meets_gate(setup):
    """precision 3 (no 8-bit products anywhere) against the same oracle outputs, and how far the default mode is from it."""
    w, rgb_a, rgb_q, tokens, ref, model2 = setup
    from oryon_b200 import _lib
    emb = model2.encode_tokens(tokens[0])[None].expand(2, -1, -1).contiguous()
    out2 = {k: v.cpu() for k, v in model2.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb).items()}
    _lib.destroy_all()  # one model per handle
    model = Oryon(None, "cuda:0", state_dict=w, precision=3)
    emb = model.encode_tokens(tokens[0])[None].expand(2, -1, -1).contiguous()
    out = model.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb)
    keys = ("featmap_a", "featmap_q", "mask_a", "mask_q")
    errs = {k: _maxerr(out[k], ref[k]) for k in keys}
    errs2 = {k: _maxerr(out2[k], ref[k]) for k in keys}
    gap = {k: _maxerr(out[k], out2[k]) for k in keys}
    print("precision 3 vs oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    print("precision 2 vs oracle:", {k: f"{v:.2e}" for k, v in errs2.items()})
    print("precision 2 vs precision 3:", {k: f"{v:.2e}" for k, v in gap.items()})
    assert all(v < TOL for v in errs.values()), errs
    assert all(v < TOL for v in errs2.values()), errs2
    _lib.destroy_all()


def test_single_pass_precision_is_close(setup):
    w, rgb_a, rgb_q, tokens, ref, _ = setup
    from oryon_b200 import _lib
    _lib.destroy_all()  # one model per handle: reload with precision 1
    model = Oryon(None, "cuda:0", state_dict=w, precision=1)
    emb = model.encode_tokens(tokens[0])[None].expand(2, -1, -1).contiguous()
    out = model.forward_tensors(rgb_a.cuda(), rgb_q.cuda(), emb)
    err = max(_maxerr(out[k], ref[k]) for k in ("featmap_a", "featmap_q", "mask_a", "mask_q"))
    print(f"precision 1 max abs err {err:.3e}")
    assert err < 5e-2
    _lib.destroy_all()
