"""N4 checkpoint ingest (oryon_b200/checkpoint.py) against the reference's own key rewriting (net.py:99-139, recorded by
oracle/make_golden_ckpt.py) and torchvision's truncated swin_b (net.py:45-58)."""
import json
import os

import torch

from oryon_b200 import checkpoint as ck
from oryon_b200 import synth_backbone as sb

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_catseg_key_rewriting_matches_reference():
    gold = json.load(open(os.path.join(GOLDEN, "ckpt_remap.json")))
    model = {k: i for i, k in enumerate(gold["keys"])}
    assert ck.remap_catseg_state_dict(model, "clip") == gold["clip"]
    assert ck.remap_catseg_state_dict(model, "dino") == gold["other_vlm"]
    assert list(ck.remap_catseg_state_dict(model, "clip")) == list(gold["clip"])      # same insertion order


def test_swin_truncation_matches_feature_extractor():
    from torchvision.models import swin_b
    from torchvision.models.feature_extraction import create_feature_extractor
    swin = swin_b(weights=None)
    nodes = {"features.1.1.add_1": "guidance3", "features.2.reduction": "guidance2", "features.4.reduction": "guidance1"}
    kept = create_feature_extractor(swin, return_nodes=nodes).state_dict()
    got = ck.swin_state_dict(swin.state_dict())
    float_kept = {"guidance_backbone." + k for k, v in kept.items() if torch.is_floating_point(v)}
    assert {k for k, v in got.items() if torch.is_floating_point(v)} == float_kept
    # and it is exactly the guidance part of the network's own state_dict
    ours = {k for k in sb.oryon_state_dict(11) if k.startswith("guidance_backbone.")}
    assert ours <= set(got) and all(k in ours for k in float_kept)


def test_assemble_order_and_lightning_prefix(tmp_path):
    sd = sb.oryon_state_dict(11)
    fusion_key, dec_key, clip_key = "fusion.conv1.weight", "decoder.head.weight", "vlm.clip_model.visual.conv1.weight"
    catseg = {"sem_seg_head.predictor.transformer.conv1.weight": sd[fusion_key] + 1,
              "sem_seg_head.predictor.transformer.head.weight": sd[dec_key] + 1,
              "sem_seg_head.predictor.clip_model.visual.conv1.weight": sd[clip_key] + 1,
              "backbone.something": torch.zeros(1)}
    light = {"state_dict": {"model." + dec_key: sd[dec_key] + 2, "feature_loss.w": torch.zeros(1), "pointdsc_solver.sigma": torch.ones(1)}}
    path = tmp_path / "last.ckpt"
    torch.save(light, path)
    clip = {k: v for k, v in sd.items() if k.startswith("vlm.")}
    swin = {k: v for k, v in sd.items() if k.startswith("guidance_backbone.")}
    out = ck.assemble_state_dict(clip, swin, catseg, ck.lightning_model_state_dict(str(path)))
    assert torch.equal(out[fusion_key], sd[fusion_key] + 1)          # CATSeg fills fusion
    assert torch.equal(out[clip_key], sd[clip_key] + 1)              # ... and overrides the stock CLIP (net.py:127-134)
    assert torch.equal(out[dec_key], sd[dec_key] + 2)                # the Lightning checkpoint wins last
    assert "backbone.something" not in out and not any(k.startswith(("feature_loss", "pointdsc")) for k in out)
    assert set(out) == set(clip) | set(swin) | {fusion_key, dec_key}


def test_clip_archive_loader(tmp_path):
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = torch.nn.Linear(4, 4).half()
            self.register_buffer("input_resolution", torch.tensor(336))
            self.register_buffer("context_length", torch.tensor(77))
            self.register_buffer("vocab_size", torch.tensor(49408))

        def forward(self, x):
            return self.visual(x)

    path = str(tmp_path / "ViT-tiny.pt")
    torch.jit.save(torch.jit.script(Tiny()), path)
    sd = ck.clip_state_dict(path)
    assert set(sd) == {"vlm.clip_model.visual.weight", "vlm.clip_model.visual.bias"}
    assert all(v.dtype == torch.float32 for v in sd.values())
    plain = str(tmp_path / "plain.pt")
    torch.save({"visual.weight": torch.zeros(2, 2, dtype=torch.float16)}, plain)
    assert ck.clip_state_dict(plain)["vlm.clip_model.visual.weight"].dtype == torch.float32


def test_reference_layout_discovery(tmp_path):
    """``Oryon(args, device)`` without an explicit state_dict reads the files of a reference installation (net.py:27-34, :99-139,
    run_test.py:42): paths, the use_catseg_ckpt / eval.ckpt switches, the load order, and a loud error naming what is missing."""
    import pytest
    home, root = tmp_path / "home", tmp_path / "repo"
    args = dict(model=dict(use_catseg_ckpt=True, image_encoder=dict(vlm="clip")), eval=dict(ckpt="ckpts/last.ckpt"))
    files = ck.reference_layout_files(args, str(root), str(home))
    assert files == {"clip": str(home / ".cache/clip/ViT-L-14-336px.pt"), "swin": str(home / ".cache/torch/hub/checkpoints/swin_b-68c6b09e.pth"),
                     "catseg": str(root / "pretrained_models/catseg.pth"), "lightning": str(root / "ckpts/last.ckpt")}
    assert ck.reference_layout_files(dict(model=dict(use_catseg_ckpt=False)), str(root), str(home))["catseg"] is None
    assert ck.reference_layout_files(None, str(root), str(home))["lightning"] is None
    with pytest.raises(FileNotFoundError, match="ViT-L-14-336px.pt"):
        ck.reference_layout_state_dict(args, str(root), str(home))
    for f in files.values():
        os.makedirs(os.path.dirname(f), exist_ok=True)
    torch.save({"visual.conv1.weight": torch.ones(2, 2, dtype=torch.float16)}, files["clip"])
    from torchvision.models import swin_b
    torch.save(swin_b(weights=None).state_dict(), files["swin"])
    torch.save({"model": {"sem_seg_head.predictor.clip_model.visual.conv1.weight": torch.full((2, 2), 2.0),
                          "sem_seg_head.predictor.transformer.head.weight": torch.full((1,), 3.0)}}, files["catseg"])
    with pytest.raises(FileNotFoundError, match="last.ckpt"):
        ck.reference_layout_state_dict(args, str(root), str(home))
    torch.save({"state_dict": {"model.decoder.head.weight": torch.full((1,), 4.0)}}, files["lightning"])
    sd = ck.reference_layout_state_dict(args, str(root), str(home))
    assert float(sd["vlm.clip_model.visual.conv1.weight"][0, 0]) == 2.0          # CATSeg's CLIP over the stock download
    assert float(sd["decoder.head.weight"][0]) == 4.0                              # the Lightning checkpoint last
    assert any(k.startswith("guidance_backbone.features.") for k in sd)
