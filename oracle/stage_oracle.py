"""CPU restatement of the reference's per-sample staging between the decoded frame and the network input.
TEST INFRASTRUCTURE (checker for ``oryon_stage_inputs``), pinned by tests/golden/stage_0.npz.

  rgb   utils/data/common.py:48-49 (``rgb.transpose(2,0,1) / 255.`` -> float64), utils/augmentations.py:137 under torchvision
        0.13 (``F.interpolate(mode='bilinear', align_corners=False)`` on the float64 tensor), datasets.py:205 (``.to(float32)``)
  mask  common.py:62-64 (``where(mask == mask_id, 1, 0)``), augmentations.py:138 (nearest, through float32), datasets.py:207
"""
import numpy as np
import torch
import torch.nn.functional as F


def stage_rgb(rgb_u8: np.ndarray, size) -> torch.Tensor:
    x = torch.tensor(rgb_u8.transpose(2, 0, 1) / 255.)
    return F.interpolate(x[None], size=list(size), mode="bilinear", align_corners=False)[0].to(torch.float32)


def stage_mask(mask: np.ndarray, mask_id: int, size) -> torch.Tensor:
    m = torch.where(torch.tensor(mask) == mask_id, 1, 0)
    return F.interpolate(m[None, None].to(torch.float32), size=list(size), mode="nearest")[0, 0].to(torch.int64).to(torch.uint8)
