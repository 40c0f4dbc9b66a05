"""Batch staging for the inference path: what the reference does per sample on DataLoader workers
between the decoded frame and the network input -- ``preprocess_item`` (utils/data/common.py:40-71), the test-time
``resize`` (utils/augmentations.py:129-164) and ``CollateWrapper`` (datasets.py:138-245) -- done once per batch on the GPU
(``oryon_stage_inputs``), from pinned host buffers.

``GpuCollate`` returns a batch dict with the reference's schema (datasets.py:202-245: ``anchor`` / ``query`` views with
``rgb [B,3,224,224] f32``, ``mask [B,224,224] u8``, ``orig_depth``, ``camera``, ``pose``, ``sizes``, ``instance_id``; top level
``prompt``, ``instance_id``, ``cls_id``, ``valid``) with two deliberate layout changes that ``FPM_Pipeline`` understands:
``rgb`` / ``mask`` already live on the GPU, and ``orig_depth`` is ONE stacked device tensor ``[B,H,W]`` instead of a list
(frames of one batch share a size in NOCS / TOYL), so lifting needs no per-pair copies.
Training-only entries (``corrs``, ``all_corrs``, resized ``depth``, ``box``, ``orig_rgb``) are not produced.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._torch_glue import as_device, ptr, require_cuda, stream_ptr


def stage_inputs(rgb_u8: Tensor, mask: Optional[Tensor], mask_ids: Optional[Tensor], size: Sequence[int] = (224, 224), device=None):
    """``rgb_u8 [B,H,W,3] uint8`` (+ ``mask [B,H,W]`` uint8 / int32 label images and the object's label per frame) ->
    ``(rgb float32 [B,3,h,w] in [0,1], mask uint8 [B,h,w] in {0,1} or None)`` on the GPU."""
    dev = torch.device(device) if device is not None else (rgb_u8.device if rgb_u8.is_cuda else require_cuda())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if rgb_u8.dtype != torch.uint8 or rgb_u8.dim() != 4 or rgb_u8.shape[3] != 3:
        raise ValueError("stage_inputs: rgb must be uint8 [B,H,W,3]")
    B, H, W, _ = rgb_u8.shape
    rgb = as_device(rgb_u8, dev)
    m, ids, out_mask = None, None, None
    if mask is not None:
        if mask.dtype not in (torch.uint8, torch.int32):
            mask = mask.to(torch.int32)
        if mask.shape != (B, H, W):
            raise ValueError("stage_inputs: mask must be [B,H,W] like the frames")
        m = as_device(mask, dev)
        ids = as_device(mask_ids, dev, torch.int32) if mask_ids is not None else None
        out_mask = torch.empty(B, size[0], size[1], dtype=torch.uint8, device=dev)
    out = torch.empty(B, 3, size[0], size[1], dtype=torch.float32, device=dev)
    _lib.check(_lib.load().oryon_stage_inputs(_lib.handle(dev.index), ptr(rgb), ptr(m), int(m is not None and m.dtype == torch.int32), ptr(ids),
                                              B, H, W, int(size[0]), int(size[1]), ptr(out), ptr(out_mask), stream_ptr(dev)))
    return out, out_mask


class GpuCollate:
    """Collate for decoded samples.  A sample is ``(item_a, item_q, prompt, pose, cls_id, instance_id, valid)`` with items
    as the dataset readers return them (utils/data/nocs.py ``get_item_data``): ``rgb`` uint8 HWC, ``mask`` label image,
    ``depth`` (mm), ``camera [3,3]``, ``instance_id``, ``metadata['mask_ids']``, ``metadata['poses']``."""

    def __init__(self, img_size: Sequence[int] = (224, 224), device=None):
        self.img_size = tuple(int(v) for v in img_size)
        self.device = device
        self._pinned: Dict[tuple, Tensor] = {}
        self._uploaded: Dict[str, torch.cuda.Event] = {}   # per view: the H2D copies that last read the pinned buffers

    def _pin(self, key: str, shape, dtype) -> Tensor:
        k = (key, tuple(shape), dtype)
        if k not in self._pinned:
            self._pinned[k] = torch.empty(shape, dtype=dtype).pin_memory()
        return self._pinned[k]

    def _view(self, tag: str, items: List[dict]) -> dict:
        B = len(items)
        H, W = items[0]["mask"].shape
        if any(it["mask"].shape != (H, W) for it in items):
            raise ValueError("GpuCollate: frames of one batch must share a size")
        if tag in self._uploaded:      # the previous batch's asynchronous uploads must have left the pinned buffers
            self._uploaded[tag].synchronize()
        rgb = self._pin(tag + "rgb", (B, H, W, 3), torch.uint8)
        mask = self._pin(tag + "mask", (B, H, W), torch.uint8 if np.asarray(items[0]["mask"]).dtype == np.uint8 else torch.int32)
        depth_dtype = {np.dtype(np.uint16): torch.int32, np.dtype(np.int64): torch.int32}.get(np.asarray(items[0]["depth"]).dtype, None)
        d0 = torch.as_tensor(np.asarray(items[0]["depth"]))
        depth = self._pin(tag + "depth", (B, H, W), depth_dtype or d0.dtype)
        for b, it in enumerate(items):
            rgb[b].copy_(torch.as_tensor(np.ascontiguousarray(it["rgb"])))
            mask[b].copy_(torch.as_tensor(np.ascontiguousarray(it["mask"])))
            depth[b].copy_(torch.as_tensor(np.ascontiguousarray(np.asarray(it["depth"]).reshape(H, W))))
        ids = torch.tensor([int(it["metadata"]["mask_ids"][0]) for it in items], dtype=torch.int32)
        dev = self.device
        rgb_f, mask_u8 = stage_inputs(rgb, mask, ids, self.img_size, dev)
        depth_dev = as_device(depth, rgb_f.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(rgb_f.device))
        self._uploaded[tag] = ev
        return dict(rgb=rgb_f, mask=mask_u8, orig_depth=depth_dev, eval_depth=depth.clone(),
                    camera=torch.stack([torch.as_tensor(np.asarray(it["camera"], dtype=np.float64)).reshape(3, 3) for it in items]),
                    pose=torch.stack([torch.as_tensor(np.asarray(it["metadata"]["poses"][0], dtype=np.float64)) for it in items]),
                    sizes=torch.tensor([[H, W]] * B), instance_id=[it["instance_id"] for it in items])

    def __call__(self, data: Sequence[tuple]) -> dict:
        items_a, items_q, prompts, poses, cls_ids, ids, valids = [], [], [], [], [], [], []
        for item_a, item_q, prompt, pose, cls_id, instance_id, valid in data:
            items_a.append(item_a), items_q.append(item_q), prompts.append(prompt), cls_ids.append(cls_id), ids.append(instance_id)
            valids.append(1. if valid else 0.)
            if pose is not None:
                poses.append(pose)
        out = dict(anchor=self._view("a", items_a), query=self._view("q", items_q), prompt=prompts, valid=torch.tensor(valids),
                   instance_id=ids, cls_id=cls_ids)
        if poses:
            out["pose"] = torch.tensor(np.stack(poses, axis=0))
        return out
