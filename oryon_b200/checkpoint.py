"""Checkpoint ingest for the inference path: builds the ``Oryon.state_dict()``-keyed dictionary that
``oryon_b200.net.Oryon.load_state_dict`` hands to the library, from the files the reference reads.

The reference fills its network in four steps; later steps overwrite earlier ones:

1. OpenAI CLIP ``ViT-L/14@336px`` through ``clip.load(..., jit=False)`` and ``.to(torch.float32)`` (models/vlm.py:19-22)
   -> keys ``vlm.clip_model.*``
2. torchvision ``swin_b(weights=DEFAULT)`` truncated by ``create_feature_extractor`` (net.py:45-58)
   -> keys ``guidance_backbone.features.*``
3. ``pretrained_models/catseg.pth`` with the prefix rewriting of ``Oryon.init_all`` (net.py:99-139), non-strict
4. the Lightning checkpoint given to ``trainer.test(..., ckpt_path=args.eval.ckpt)`` (run_test.py:42): ``state_dict``
   entries ``model.*`` are the network's.

Host-only plumbing (key strings and tensors; no arithmetic, no GPU).  The rewriting rules are pinned by
tests/golden/ckpt_remap.json, recorded by executing the reference's own loop (oracle/make_golden_ckpt.py).
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional

import torch
from torch import Tensor

CATSEG_FUSION_PREFIX = "sem_seg_head.predictor.transformer"
CATSEG_CLIP_PREFIX = "sem_seg_head.predictor.clip_model"


def remap_catseg_state_dict(ckpt_model: Mapping[str, Tensor], vlm: str = "clip") -> Dict[str, Tensor]:
    """``ckpt['model']`` of CATSeg -> the entries of Oryon's state_dict it provides (net.py:99-139).

    Every ``str.replace`` of the reference acts on ALL occurrences of the pattern, not only on the prefix, and the
    second and third rules are tested on the already rewritten key; both are kept."""
    out: Dict[str, Tensor] = {}
    for key, value in ckpt_model.items():
        if not key.startswith(CATSEG_FUSION_PREFIX):
            continue
        new = key.replace(CATSEG_FUSION_PREFIX, "fusion")
        if new.startswith("fusion.decoder"):
            new = new.replace("fusion.decoder", "decoder.decoder")
        if new.startswith("fusion.head"):
            new = new.replace("fusion.head", "decoder.head")
        out[new] = value
    if vlm == "clip":
        for key, value in ckpt_model.items():
            if key.startswith(CATSEG_CLIP_PREFIX):
                out[key.replace(CATSEG_CLIP_PREFIX, "vlm.clip_model")] = value
    return out


def clip_state_dict(path: str) -> Dict[str, Tensor]:
    """The OpenAI CLIP download (a TorchScript archive, or a plain state_dict file) as ``vlm.clip_model.*`` float32
    entries -- what ``clip.load(name, device, jit=False)`` + ``.to(torch.float32)`` leaves in the module
    (models/vlm.py:19-22).  Non-parameter entries of the archive (``input_resolution``, ``context_length``,
    ``vocab_size``), which ``clip.model.build_model`` deletes, are dropped."""
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd.state_dict()
    out = {}
    for k, v in sd.items():
        if k in ("input_resolution", "context_length", "vocab_size"):
            continue
        out["vlm.clip_model." + k] = v.to(torch.float32) if torch.is_floating_point(v) else v
    return out


def swin_state_dict(path_or_sd) -> Dict[str, Tensor]:
    """torchvision ``swin_b`` weights -> ``guidance_backbone.features.{0..4}.*``: the part of the model that feeds the
    three return nodes of net.py:49-53 (``features.1.1.add_1``, ``features.2.reduction``, ``features.4.reduction``);
    ``create_feature_extractor`` drops everything after them."""
    sd = torch.load(path_or_sd, map_location="cpu") if isinstance(path_or_sd, str) else path_or_sd
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if parts[0] != "features" or int(parts[1]) > 4:
            continue
        if parts[1] == "4" and parts[2] != "reduction" and parts[2] != "norm":
            continue
        out["guidance_backbone." + k] = v
    return out


def lightning_model_state_dict(ckpt) -> Dict[str, Tensor]:
    """``model.*`` entries of a Lightning checkpoint (path or loaded dict) with the prefix stripped: the state of
    ``FPM_Pipeline.model`` (pipeline.py:73); other entries (losses, PointDSC solver, ...) are ignored."""
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd = ckpt.get("state_dict", ckpt)
    return {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}


def assemble_state_dict(clip: Optional[Mapping[str, Tensor]] = None, swin: Optional[Mapping[str, Tensor]] = None,
                        catseg_model: Optional[Mapping[str, Tensor]] = None, lightning: Optional[Mapping[str, Tensor]] = None,
                        vlm: str = "clip") -> Dict[str, Tensor]:
    """Merge in the reference's load order (module docstring).  Shape-mismatching or unknown entries are not
    filtered here: ``Oryon.load_state_dict`` ignores names the network does not have, like ``strict=False``."""
    out: Dict[str, Tensor] = {}
    for part in (clip, swin, remap_catseg_state_dict(catseg_model, vlm) if catseg_model is not None else None, lightning):
        if part:
            out.update(part)
    return out


# the files the reference's constructors read, where its libraries put them
CLIP_CACHE_FILE = os.path.join("~", ".cache", "clip", "ViT-L-14-336px.pt")                  # clip.load("ViT-L/14@336px"), models/vlm.py:19
SWIN_HUB_FILE = os.path.join("~", ".cache", "torch", "hub", "checkpoints", "swin_b-68c6b09e.pth")   # swin_b(weights=Swin_B_Weights.DEFAULT), net.py:46
CATSEG_FILE = os.path.join("pretrained_models", "catseg.pth")                               # net.py:104


def reference_layout_files(args=None, root: str = ".", home: Optional[str] = None) -> Dict[str, Optional[str]]:
    """Where ``Oryon(args, device)`` finds its weights in a reference installation: the OpenAI CLIP download in ``~/.cache/clip``,
    torchvision's ``swin_b`` weights in the torch hub cache, ``pretrained_models/catseg.pth`` when ``args.model.use_catseg_ckpt``
    (net.py:102-104) and the Lightning checkpoint ``args.eval.ckpt`` that ``trainer.test(..., ckpt_path=...)`` loads on top
    (run_test.py:42).  ``None`` for a part the configuration does not ask for."""
    def get(path, default=None):
        cur = args
        for key in path.split("."):
            if cur is None:
                return default
            cur = cur.get(key) if isinstance(cur, Mapping) else getattr(cur, key, None)
        return default if cur is None else cur

    expand = (lambda p: p.replace("~", home, 1)) if home is not None else os.path.expanduser
    ckpt = get("eval.ckpt")
    return {"clip": expand(CLIP_CACHE_FILE), "swin": expand(SWIN_HUB_FILE),
            "catseg": os.path.join(root, CATSEG_FILE) if get("model.use_catseg_ckpt", False) else None,
            "lightning": (ckpt if os.path.isabs(str(ckpt)) else os.path.join(root, str(ckpt))) if ckpt else None}


def reference_layout_state_dict(args=None, root: str = ".", home: Optional[str] = None) -> Dict[str, Tensor]:
    """The state_dict the reference ends up with after ``Oryon.__init__`` and the checkpoint load, read from the files of
    ``reference_layout_files``; a file the configuration asks for and that is missing raises ``FileNotFoundError`` naming it."""
    files = reference_layout_files(args, root, home)
    missing = [f"{k}: {v}" for k, v in files.items() if v is not None and not os.path.exists(v)]
    if missing:
        raise FileNotFoundError("pretrained weights not found (" + "; ".join(missing) + ")")
    vlm = "clip"
    return assemble_state_dict(clip_state_dict(files["clip"]), swin_state_dict(files["swin"]),
                               torch.load(files["catseg"], map_location="cpu")["model"] if files["catseg"] else None,
                               lightning_model_state_dict(files["lightning"]) if files["lightning"] else None, vlm)
