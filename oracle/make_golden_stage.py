"""Generate ``tests/golden/stage_0.npz``: the batch the reference's loader hands to the network for two synthetic decoded
frames -- ``preprocess_item`` (utils/data/common.py:40-71), the test-time ``resize`` transform
(utils/augmentations.py:129-164) and ``CollateWrapper`` (datasets.py:138-245), all UNMODIFIED reference code.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_stage.py

Two environment notes.  (1) The reference pins torch 1.12 / torchvision 0.13 (environment.yml:180), where ``F.resize`` on a
tensor does not antialias; torchvision 0.26 here defaults to ``antialias=True``, so ``F.resize`` is wrapped to pass
``antialias=False`` -- the reference's own behaviour.  (2) ``datasets.py`` imports the dataset readers (matplotlib, open3d,
...), so ``CollateWrapper`` is exec'd from its source text as it lies under /root/reference instead of importing the module.
"""
import functools
import hashlib
import os
import sys
from typing import Sequence, Tuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

ref_shims.install()

import torchvision.transforms.functional as TVF  # noqa: E402

_orig_resize = TVF.resize
TVF.resize = functools.wraps(_orig_resize)(lambda img, size, interpolation, **kw: _orig_resize(img, size, interpolation, **{**kw, "antialias": False}))

from utils.data import common as ref_common  # noqa: E402  (reference)
from utils import augmentations as ref_augs  # noqa: E402

from oryon_b200 import synth  # noqa: E402


def reference_collate_wrapper():
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "datasets.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("class CollateWrapper"))
    end = next(i for i, l in enumerate(src) if l.startswith("class Shapenet6DDataset"))
    env = {"torch": torch, "np": np, "Sequence": Sequence, "Tuple": Tuple}
    exec("\n".join(src[start:end]), env)
    return env["CollateWrapper"]


def main(seed=0):
    frames = synth.raw_frames(seed, 2)
    items = []
    for f in frames:
        item = dict(rgb=f["rgb"].copy(), mask=f["mask"].copy(), depth=f["depth"].copy(), camera=np.asarray(synth.NOCS_INTRINSICS).reshape(3, 3),
                    instance_id=f["instance_id"], metadata=dict(mask_ids=[f["mask_id"]], poses=[np.eye(4)], cls_ids=[1], cls_names=["mug"]))
        items.append(ref_common.preprocess_item(item))
    corrs = torch.zeros(10, 4)
    a, q, corrs = ref_augs.resize((224, 224))((items[0], items[1], corrs))
    batch = reference_collate_wrapper()(500)([(a, q, ["mug"] * 81, torch.zeros(500, 4), torch.zeros(500, 4), None, 1, "pair0", True)])
    rgb = torch.cat([batch["anchor"]["rgb"], batch["query"]["rgb"]]).numpy()
    mask = torch.cat([batch["anchor"]["mask"], batch["query"]["mask"]]).numpy()
    assert rgb.dtype == np.float32 and mask.dtype == np.uint8 and rgb.shape == (2, 3, 224, 224)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"stage_{seed}.npz"),
                        rgb_rows=rgb[:, :, ::8, :], mask=mask, rgb_sha=hashlib.sha256(rgb.tobytes()).hexdigest(),
                        sizes=torch.cat([batch["anchor"]["sizes"], batch["query"]["sizes"]]).numpy(),
                        box=torch.cat([batch["anchor"]["box"], batch["query"]["box"]]).numpy(),
                        in_sum=synth.tensor_checksum(torch.from_numpy(frames[0]["rgb"])) + synth.tensor_checksum(torch.from_numpy(frames[1]["mask"])))
    print(rgb.shape, rgb.mean(), mask.sum(axis=(1, 2)), batch["anchor"]["sizes"], batch["anchor"]["box"], batch["anchor"]["orig_depth"][0].shape)


if __name__ == "__main__":
    main()
