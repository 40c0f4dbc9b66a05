"""Generate ``tests/golden/toyl_tree_0.{json,npz}``: what the reference's own TOYL readers return for the synthetic dataset
tree of ``oryon_b200.synth.write_toyl_tree`` -- ``TOYLDataset(args, eval=True)[i]`` for every pair (datasets.py:546-713, with
``utils/data/toyl.py`` and ``utils/data/common.py`` underneath) and ``get_object_info()`` -- all UNMODIFIED reference code.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_toyl.py

Environment notes.  matplotlib and pytz are stubbed as in ``make_golden_nocs.py``; ``datasets.py`` is exec'd from its source
text (from ``set_seed`` to the end, with the dataset classes this script does not use left undefined-at-call); ``F.resize`` is
wrapped with ``antialias=False``.  ``utils/data/toyl.py`` reads the object models with the third-party ``plyfile`` package
(not installed): a stand-in ``PlyData.read`` below parses the fixed property layout the synthetic writer emits (six float
vertex properties, ``list uchar int vertex_indices`` faces; ASCII via text split, binary via one structured numpy read) --
written independently of ``oryon_b200.utils.data.toyl.read_ply``, so the model arrays pin that parser too.
"""
import functools
import hashlib
import json
import os
import pickle
import sys
import tempfile
import types
from typing import Any, Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

ref_shims.install()
for name in ("pytz", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "plyfile"):
    sys.modules.setdefault(name, types.ModuleType(name))


class _PlyData:
    """Stand-in for ``plyfile.PlyData`` (see the module docstring)."""

    @staticmethod
    def read(path):
        raw = open(path, "rb").read()
        head, body = raw.split(b"end_header\n", 1)
        lines = head.decode().split("\n")
        nv = int(next(l for l in lines if l.startswith("element vertex")).split()[2])
        nf = int(next(l for l in lines if l.startswith("element face")).split()[2])
        names = ["x", "y", "z", "nx", "ny", "nz"]
        if "format ascii" in head.decode():
            rows = body.decode().split("\n")
            v = np.array([[np.float32(t) for t in r.split()] for r in rows[:nv]], dtype=np.float32)
            faces = [np.array(r.split()[1:], dtype=np.int32) for r in rows[nv:nv + nf]]
        else:
            v = np.frombuffer(body, dtype="<f4", count=nv * 6).reshape(nv, 6)
            rec = np.frombuffer(body, dtype=np.dtype([("n", "u1"), ("idx", "<i4", (3,))]), count=nf, offset=nv * 24)
            faces = [np.array(r["idx"]) for r in rec]
        fa = np.empty(nf, dtype=object)
        fa[:] = faces
        return {"vertex": {n: v[:, i].copy() for i, n in enumerate(names)}, "face": {"vertex_indices": fa}}


sys.modules["plyfile"].PlyData = _PlyData
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].collections = sys.modules["matplotlib.collections"]

import torchvision  # noqa: E402
import torchvision.transforms.functional as TVF  # noqa: E402

_orig_resize = TVF.resize
TVF.resize = functools.wraps(_orig_resize)(lambda img, size, interpolation, **kw: _orig_resize(img, size, interpolation, **{**kw, "antialias": False}))

from omegaconf import DictConfig  # noqa: E402  (shim: dict with attribute access)
from utils.data import common as ref_common, toyl as ref_toyl  # noqa: E402  (reference)
from utils import augmentations as ref_augs  # noqa: E402
from utils.misc import torch_sample_select, unique_matches  # noqa: E402

from oryon_b200 import synth  # noqa: E402


def reference_dataset_classes():
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "datasets.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def set_seed"))
    end = len(src)
    from PIL import Image
    from torch.nn.functional import interpolate
    from torch.utils.data import Dataset
    env = {"os": os, "json": json, "pickle": pickle, "torch": torch, "np": np, "Dataset": Dataset, "DictConfig": DictConfig, "Any": Any,
           "Dict": Dict, "Tuple": Tuple, "Sequence": Sequence, "Union": Union, "List": List, "join": os.path.join, "torchvision": torchvision,
           "interpolate": interpolate, "toyl": ref_toyl, "common": ref_common, "torch_sample_select": torch_sample_select,
           "unique_matches": unique_matches, "Tensor": torch.Tensor, "Image": Image}
    env.update({k: getattr(ref_augs, k) for k in dir(ref_augs) if not k.startswith("_")})      # from utils.augmentations import *
    exec("\n".join(src[start:end]), env)
    return env


def cfg(d):
    return DictConfig({k: cfg(v) if isinstance(v, dict) else v for k, v in d.items()})


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def view_record(item: dict) -> dict:
    rgb_u8 = torch.round(item["orig_rgb"] * 255.).to(torch.uint8).permute(1, 2, 0).numpy()      # preprocess_item: rgb.transpose(2,0,1) / 255.
    md = item["metadata"]
    return dict(instance_id=item["instance_id"], hw_size=[int(v) for v in item["hw_size"]], mask_ids=[int(v) for v in md["mask_ids"]],
                cls_ids=[int(v) for v in md["cls_ids"]], cls_names=list(md["cls_names"]), cls_descs=[list(d) for d in md["cls_descs"]],
                n_poses=len(md["poses"]), pose0=md["poses"][0].numpy().tolist(), camera=np.asarray(item["camera"]).tolist(),
                rgb_sha=sha(rgb_u8), depth_sha=sha(item["orig_depth"].numpy().astype(np.int64)),
                mask224_sha=sha(item["mask"].numpy().astype(np.uint8)), mask224_sum=int(item["mask"].sum()))


def main(seed=0):
    env = reference_dataset_classes()
    out_json, out_npz = {}, {}
    with tempfile.TemporaryDirectory() as d:
        info = synth.write_toyl_tree(d, seed)
        for obj_split, mask in (("all", "predicted"), ("cars", "oracle")):
            args = cfg(dict(augs=dict(), debug_valid="no", use_seed=False, seed=1,
                            dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj=obj_split)),
                            test=dict(mask=mask, add_description="yes" if obj_split == "all" else "wrong")))
            ds = env["TOYLDataset"](args, eval=True)
            recs = []
            for i in range(len(ds)):
                item_a, item_q, prompt, _, orig_corrs, pose, obj_id, instance_id, valid = ds[i]
                recs.append(dict(instance_id=instance_id, obj_id=int(obj_id), valid=bool(valid), prompt=list(prompt), pose=np.asarray(pose).tolist(),
                                 n_corrs=int(orig_corrs.shape[0]), anchor=view_record(item_a), query=view_record(item_q)))
            out_json[obj_split] = dict(length=len(ds), tracked=list(ds.tracked_instances), samples=recs)
        models, diams, symms = ds.get_object_info()
        out_json["objects"] = {str(k): dict(diameter=float(diams[k]), n_symmetries=len(symms[k])) for k in models}
        for k in models:
            out_npz[f"{k}/pts"], out_npz[f"{k}/normals"], out_npz[f"{k}/faces"] = models[k]["pts"], models[k]["normals"], models[k]["faces"]
            out_npz[f"{k}/sym_R"] = np.stack([s["R"] for s in symms[k]])
            out_npz[f"{k}/sym_t"] = np.stack([s["t"] for s in symms[k]])
    gold = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gold, f"toyl_tree_{seed}.json"), "w") as f:
        json.dump(out_json, f, indent=1)
    np.savez_compressed(os.path.join(gold, f"toyl_tree_{seed}.npz"), **out_npz)
    print("wrote", {k: v["length"] for k, v in out_json.items() if "length" in v}, sorted(out_json["objects"]))


if __name__ == "__main__":
    main()
