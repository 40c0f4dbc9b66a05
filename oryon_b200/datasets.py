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

import json
import os
import pickle
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


# ------------------------------------------------------------------------------------------------
# dataset reader (test time): the reference's NOCSDataset (datasets.py:369-543) producing GpuCollate samples
# ------------------------------------------------------------------------------------------------
def _cfg(obj, path: str, default=None):
    """``args.a.b`` on attribute- or mapping-style configs (hydra's DictConfig in the reference, plain dicts here)."""
    cur = obj
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


def get_mask_type(mask: str, eval: bool) -> str:
    """Which mask FILE a sample carries (datasets.py:27-45): when the mask is predicted by the network the oracle mask
    is loaded as ground truth; outside evaluation always the oracle."""
    if eval:
        return "oracle" if mask == "predicted" else mask
    return "oracle"


def nearest_resized_mask_nonempty(mask01: np.ndarray, size: Sequence[int]) -> bool:
    """``check_validity`` (utils/data/common.py:104-111) of the mask AFTER the test-time nearest resize
    (utils/augmentations.py:138): whether any sampled source pixel ``floor(dst * float32(in / out))`` is set."""
    H, W = mask01.shape
    ys = np.minimum(np.floor(np.arange(size[0], dtype=np.float32) * np.float32(H / size[0])).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(size[1], dtype=np.float32) * np.float32(W / size[1])).astype(np.int64), W - 1)
    return bool(mask01[np.ix_(ys, xs)].any())


class _PairSplitDataset:
    """Common part of the reference's test-time dataset classes (datasets.py:369-543 NOCS, :546-713 TOYL): a fixed split of
    (anchor frame, query frame, object) pairs with relative ground-truth poses.

    ``dataset[i]`` is the ``GpuCollate`` sample ``(item_a, item_q, prompt, pose, obj_id, instance_id, valid)``: the two
    decoded frames as the dataset's ``get_item_data`` returns them plus the intrinsics, the 1 + len(templates) prompt strings
    (:515-532 / :685-702), the relative ground-truth pose with its translation in metres (:441-443), the object key, the pair
    id and the validity flag (:492-497: both resized masks non-empty and ground-truth correspondences present).  What the
    reference does next on the DataLoader workers (``preprocess_item``, resize, ``CollateWrapper``) is ``GpuCollate``'s job,
    on the GPU.  Training-only products (sampled correspondences, augmentations) are not produced."""

    def __init__(self, args, eval: bool = True):
        self.eval = eval
        self.root = _cfg(args, "dataset.root")
        self.max_corrs = int(_cfg(args, "dataset.max_corrs", 500))
        self.img_size = tuple(_cfg(args, "dataset.img_size", (224, 224)))
        self.mask_type = _cfg(args, "test.mask", "oracle")
        self.add_description = _cfg(args, "test.add_description", "yes")
        part = "test" if eval else "train"
        self.name = _cfg(args, f"dataset.{part}.name")
        self.split = _cfg(args, f"dataset.{part}.split")
        self.obj = str(_cfg(args, f"dataset.{part}.obj"))
        self.K = self._reader.get_camera() if not hasattr(self, "K") else self.K
        self.local_root = os.path.join(self.root, self.name)
        with open(os.path.join(self.local_root, "templates.json")) as f:
            self.prompt_templates = json.load(f)
        with open(os.path.join(self.local_root, "object_splits.json")) as f:
            self.obj_ids = [int(cat) for cat in json.load(f)[self.obj]]
        self.part_data = self._reader.get_part_data(self.local_root)
        self.obj_names = self._reader.get_obj_names(self.local_root)
        self.path_split = os.path.join(self.local_root, "fixed_split", self.split)
        self._obj_data = None
        with open(os.path.join(self.path_split, "instance_list.txt")) as f:
            lines = f.readlines()
        with open(os.path.join(self.path_split, "annots.pkl"), "rb") as f:
            annots = pickle.load(f)
        self.instances, self.poses, self.corrs = [], [], []
        for line in lines:
            inst = self._reader.parse_pair_line(line)
            if inst[5] in self.obj_ids:
                key = "_".join(str(e) for e in inst[1:])
                pose = np.array(annots[key]["gt"], dtype=np.float64, copy=True)
                pose[:3, 3] = pose[:3, 3] / 1000.
                self.poses.append(pose)
                self.corrs.append(annots[key]["corrs"])
                self.instances.append(inst)
        self.tracked_instances = []
        tracked = os.path.join(self.path_split, "tracked.txt")
        if os.path.exists(tracked):
            with open(tracked) as f:
                for line in f.readlines():
                    inst = self._reader.parse_pair_line(line)
                    self.tracked_instances.append("_".join(str(e) for e in inst[1:5] + (inst[-1],)))
        self.collate = GpuCollate(self.img_size, _cfg(args, "device"))

    def __len__(self) -> int:
        return len(self.instances)

    def get_item(self, scene_id: int, img_id: int, obj_id, mask_type: str = "oracle") -> dict:
        return self._reader.get_item_data(self.local_root, scene_id, img_id, self.part_data, self.obj_names, obj_id, mask_type)

    def frame_ids(self, index: int):
        """``'scene image object'`` ids of the two frames of pair ``index`` (the ids of a prediction-CSV line)."""
        inst = self.instances[index]
        return f"{inst[1]} {inst[2]} {inst[-1]}", f"{inst[3]} {inst[4]} {inst[-1]}"

    def get_item_prompt(self, item: dict) -> List[str]:
        name = item["metadata"]["cls_names"][0]
        if self.add_description == "yes":
            name = f"{item['metadata']['cls_descs'][0][0]} {name}"
        elif self.add_description == "wrong":
            name = f"{item['metadata']['cls_descs'][0][1]} {name}"
        elif self.add_description == "desconly":
            name = f"{item['metadata']['cls_descs'][0][0]} object"
        return [name] + [template.format(name) for template in self.prompt_templates]

    def __getitem__(self, index: int):
        inst = self.instances[index]
        scene_a, img_a, scene_q, img_q, obj_id = inst[1], inst[2], inst[3], inst[4], inst[-1]
        instance_id = f"{scene_a}_{img_a}_{scene_q}_{img_q}_{obj_id}"
        mask = get_mask_type(self.mask_type, self.eval)
        item_a, item_q = self.get_item(scene_a, img_a, obj_id, mask), self.get_item(scene_q, img_q, obj_id, mask)
        valid = len(self.corrs[index]) > 0
        for item in (item_a, item_q):
            if len(item["metadata"]["mask_ids"]) != 1:      # the assertion of preprocess_item (utils/data/common.py:45)
                raise AssertionError(f" Problem with instance {item['instance_id']}: no objects found. Check cls_id!")
            item["camera"] = self.K
            valid = valid and nearest_resized_mask_nonempty(np.asarray(item["mask"]) == item["metadata"]["mask_ids"][0], self.img_size)
        return item_a, item_q, self.get_item_prompt(item_a), self.poses[index], obj_id, instance_id, valid

    def get_object_info(self):
        """``(models, diameters, symmetries)`` of all objects, for ``Evaluator.add_object_info`` (datasets.py:509-513)."""
        if self._obj_data is None:
            self._obj_data = self._reader.get_obj_data(self.local_root)
        return self._obj_data


class NOCSDataset(_PairSplitDataset):
    """Test-time reader of the NOCS (REAL275) pair split in the reference's on-disk layout (datasets.py:369-457): pairs are
    ``(split, scene_a, img_a, scene_q, img_q, category id, object name)``, filtered by category id, keyed by object NAME."""
    from .utils.data import nocs as _reader

    @property
    def abs_poses(self):
        return self.part_data

    def get_obj_info(self, obj_id):
        models, diams, symms = self.get_object_info()
        return models[obj_id], diams[obj_id], symms[obj_id]


class TOYLDataset(_PairSplitDataset):
    """Test-time reader of the TOYL pair split in the reference's on-disk layout (datasets.py:546-630): pairs are
    ``(split, scene_a, img_a, scene_q, img_q, object id)``, filtered and keyed by the integer object id; the intrinsics are
    the ones the dataset class hard-codes (datasets.py:573), not ``utils.data.toyl.get_camera()``."""
    from .utils.data import toyl as _reader
    K = np.asarray([[572.4114, 0.0, 325.2611], [0.0, 573.5704, 242.0489], [0.0, 0.0, 1.0]])

    def get_obj_info(self, obj_id):
        models, diams, symms = self.get_object_info()
        return models[int(obj_id)], diams[int(obj_id)], symms[int(obj_id)]
