"""N1 on the CPU: the staging oracle (oracle/stage_oracle.py) against the batch the reference's own loader code produced
(tests/golden/stage_0.npz: preprocess_item + resize + CollateWrapper, oracle/make_golden_stage.py)."""
import hashlib
import os

import numpy as np
import torch

import stage_oracle
from oryon_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_stage_oracle_matches_reference_loader():
    g = np.load(os.path.join(GOLDEN, "stage_0.npz"))
    frames = synth.raw_frames(0, 2)
    assert str(g["in_sum"]) == synth.tensor_checksum(torch.from_numpy(frames[0]["rgb"])) + synth.tensor_checksum(torch.from_numpy(frames[1]["mask"]))
    rgb = torch.stack([stage_oracle.stage_rgb(f["rgb"], (224, 224)) for f in frames]).numpy()
    mask = torch.stack([stage_oracle.stage_mask(f["mask"], f["mask_id"], (224, 224)) for f in frames]).numpy()
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == str(g["rgb_sha"])
    assert np.array_equal(rgb[:, :, ::8, :], g["rgb_rows"])
    assert np.array_equal(mask, g["mask"]) and mask.dtype == np.uint8
