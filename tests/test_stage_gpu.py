"""N1 on the GPU: ``oryon_stage_inputs`` / ``GpuCollate`` bit-exact against the reference loader's output (golden) and the
oracle, and the staged batch through ``FPM_Pipeline``'s batched tail."""
import hashlib
import os

import numpy as np
import pytest
import torch

import stage_oracle
from gpu_util import need_gpu
from oryon_b200 import synth
from oryon_b200.datasets import GpuCollate, stage_inputs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_stage_inputs_bit_exact_vs_reference_loader():
    need_gpu()
    g = np.load(os.path.join(GOLDEN, "stage_0.npz"))
    frames = synth.raw_frames(0, 2)
    rgb_u8 = torch.stack([torch.from_numpy(f["rgb"]) for f in frames]).pin_memory()
    mask = torch.stack([torch.from_numpy(f["mask"]) for f in frames])
    ids = torch.tensor([f["mask_id"] for f in frames], dtype=torch.int32)
    rgb, m = stage_inputs(rgb_u8, mask, ids, (224, 224), "cuda:0")
    rgb, m = rgb.cpu().numpy(), m.cpu().numpy()
    assert rgb.dtype == np.float32 and rgb.shape == (2, 3, 224, 224) and m.dtype == np.uint8
    assert np.array_equal(rgb[:, :, ::8, :], g["rgb_rows"])
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == str(g["rgb_sha"])
    assert np.array_equal(m, g["mask"])


@pytest.mark.parametrize("hw,size,mask_dtype", [((480, 640), (224, 224), torch.int32), ((200, 150), (224, 224), torch.uint8),
                                                 ((37, 53), (16, 24), torch.uint8)])
def test_stage_inputs_vs_oracle_shapes(hw, size, mask_dtype):
    """Down- and up-scaling, odd sizes, int32 label images, frames without a mask."""
    need_gpu()
    frames = synth.raw_frames(3, 3, hw)
    rgb_u8 = torch.stack([torch.from_numpy(f["rgb"]) for f in frames])
    mask = torch.stack([torch.from_numpy(f["mask"]) for f in frames]).to(mask_dtype)
    ids = torch.tensor([3, 4, 5], dtype=torch.int32)
    rgb, m = stage_inputs(rgb_u8, mask, ids, size, "cuda:0")
    for b, f in enumerate(frames):
        assert torch.equal(rgb[b].cpu(), stage_oracle.stage_rgb(f["rgb"], size))
        assert torch.equal(m[b].cpu(), stage_oracle.stage_mask(f["mask"], int(ids[b]), size))
    rgb2, none = stage_inputs(rgb_u8, None, None, size, "cuda:0")
    assert none is None and torch.equal(rgb2, rgb)


def test_gpu_collate_schema():
    need_gpu()
    frames = synth.raw_frames(5, 4)
    K = np.asarray(synth.NOCS_INTRINSICS).reshape(3, 3)

    def item(f):
        return dict(rgb=f["rgb"], mask=f["mask"], depth=f["depth"], camera=K, instance_id=f["instance_id"],
                    metadata=dict(mask_ids=[f["mask_id"]], poses=[np.eye(4)]))

    data = [(item(frames[0]), item(frames[1]), ["mug"] * 81, np.eye(4), 1, "p0", True),
            (item(frames[2]), item(frames[3]), ["can"] * 81, np.eye(4), 2, "p1", False)]
    batch = GpuCollate((224, 224), "cuda:0")(data)
    a = batch["anchor"]
    assert a["rgb"].shape == (2, 3, 224, 224) and a["rgb"].is_cuda and a["mask"].dtype == torch.uint8
    assert a["orig_depth"].shape == (2, 480, 640) and a["orig_depth"].is_cuda and a["orig_depth"].dtype == torch.int32
    assert torch.equal(a["orig_depth"][1].cpu(), torch.from_numpy(frames[2]["depth"]))
    assert a["camera"].shape == (2, 3, 3) and a["pose"].shape == (2, 4, 4) and a["sizes"].tolist() == [[480, 640]] * 2
    assert batch["valid"].tolist() == [1.0, 0.0] and batch["cls_id"] == [1, 2] and batch["prompt"][1][0] == "can"
    assert torch.equal(batch["query"]["rgb"][0].cpu(), stage_oracle.stage_rgb(frames[1]["rgb"], (224, 224)))
