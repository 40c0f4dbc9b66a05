"""Generate ``tests/golden/nocs_tree_0.{json,npz}``: what the reference's own NOCS readers return for the synthetic dataset
tree of ``oryon_b200.synth.write_nocs_tree`` -- ``NOCSDataset(args, eval=True)[i]`` for every pair (datasets.py:369-543, with
``utils/data/nocs.py`` and ``utils/data/common.py`` underneath) and ``get_object_info()`` -- all UNMODIFIED reference code.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_nocs.py

Environment notes.  ``utils/data/nocs.py`` imports matplotlib and ``bop_toolkit_lib/misc.py`` imports pytz, neither installed
and neither used by the readers: both are stubbed with empty modules.  ``datasets.py`` imports every dataset reader and the
visualisation module (open3d, ...), so its source from ``set_seed`` up to ``class TOYLDataset`` is exec'd as it lies under
/root/reference instead of importing the module.  ``F.resize`` is wrapped with ``antialias=False`` as in
``make_golden_stage.py`` (the reference's pinned torchvision 0.13 does not antialias tensors).
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
for name in ("pytz", "matplotlib", "matplotlib.pyplot", "matplotlib.collections"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].collections = sys.modules["matplotlib.collections"]

import torchvision  # noqa: E402
import torchvision.transforms.functional as TVF  # noqa: E402

_orig_resize = TVF.resize
TVF.resize = functools.wraps(_orig_resize)(lambda img, size, interpolation, **kw: _orig_resize(img, size, interpolation, **{**kw, "antialias": False}))

from omegaconf import DictConfig  # noqa: E402  (shim: dict with attribute access)
from utils.data import common as ref_common, nocs as ref_nocs  # noqa: E402  (reference)
from utils import augmentations as ref_augs  # noqa: E402
from utils.misc import torch_sample_select, unique_matches  # noqa: E402

from oryon_b200 import synth  # noqa: E402


def reference_dataset_classes():
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "datasets.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def set_seed"))
    end = next(i for i, l in enumerate(src) if l.startswith("class TOYLDataset"))
    from PIL import Image
    from torch.nn.functional import interpolate
    from torch.utils.data import Dataset
    env = {"os": os, "json": json, "pickle": pickle, "torch": torch, "np": np, "Dataset": Dataset, "DictConfig": DictConfig, "Any": Any,
           "Dict": Dict, "Tuple": Tuple, "Sequence": Sequence, "Union": Union, "List": List, "join": os.path.join, "torchvision": torchvision,
           "interpolate": interpolate, "nocs": ref_nocs, "common": ref_common, "torch_sample_select": torch_sample_select,
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
        info = synth.write_nocs_tree(d, seed)
        for obj_split, mask in (("all", "predicted"), ("mugs", "oracle")):
            args = cfg(dict(augs=dict(), debug_valid="no", use_seed=False, seed=1,
                            dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj=obj_split)),
                            test=dict(mask=mask, add_description="yes" if obj_split == "all" else "wrong")))
            ds = env["NOCSDataset"](args, eval=True)
            recs = []
            for i in range(len(ds)):
                item_a, item_q, prompt, _, orig_corrs, pose, obj_id, instance_id, valid = ds[i]
                recs.append(dict(instance_id=instance_id, obj_id=obj_id, valid=bool(valid), prompt=list(prompt), pose=np.asarray(pose).tolist(),
                                 n_corrs=int(orig_corrs.shape[0]), anchor=view_record(item_a), query=view_record(item_q)))
            out_json[obj_split] = dict(length=len(ds), tracked=list(ds.tracked_instances), samples=recs)
        models, diams, symms = ds.get_object_info()
        out_json["objects"] = {k: dict(diameter=float(diams[k]), n_symmetries=len(symms[k])) for k in models}
        for k in models:
            out_npz[f"{k}/pts"], out_npz[f"{k}/normals"], out_npz[f"{k}/faces"] = models[k]["pts"], models[k]["normals"], models[k]["faces"]
            out_npz[f"{k}/sym_R"] = np.stack([s["R"] for s in symms[k]])
            out_npz[f"{k}/sym_t"] = np.stack([s["t"] for s in symms[k]])
    gold = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gold, f"nocs_tree_{seed}.json"), "w") as f:
        json.dump(out_json, f, indent=1)
    np.savez_compressed(os.path.join(gold, f"nocs_tree_{seed}.npz"), **out_npz)
    print("wrote", {k: v["length"] for k, v in out_json.items() if "length" in v}, sorted(out_json["objects"]))


if __name__ == "__main__":
    main()
