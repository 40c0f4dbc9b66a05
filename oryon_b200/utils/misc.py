"""Mirror of the helpers of the reference ``utils/misc.py`` that the inference path touches:
``torch_sample_select`` (:242-254), ``rescale_coords`` (:93-122) and ``set_deterministic_seed`` (:186-196).  Host-side index /
seed helpers: nothing here is arithmetic of the hot path (that lives in liboryon_b200.so)."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from .pcd import torch_sample_select  # noqa: F401  (same function, defined next to its callers)


def rescale_coords(coords: Tensor, orig_scale: Tuple[int, int], new_scale: Tuple[int, int]) -> Tensor:
    """A copy of ``[N,2|4]`` (or batched ``[B,N,2|4]``) YX coordinates / correspondences moved from ``orig_scale`` to ``new_scale``
    and clamped to the new image (utils/misc.py:93-122; the test loop uses it on the correspondences of tracked pairs before
    drawing them, pipeline.py:333).  Dtype-preserving like the reference: integer coordinates are truncated by the in-place
    assignment."""
    new = coords.clone()
    squeeze = new.dim() == 2
    if squeeze:
        new = new.unsqueeze(0)
    assert new.shape[-1] in (2, 4), " works only with 2D keypoints or 2D correspondences"
    for col, axis in ((0, 0), (1, 1)) + (((2, 0), (3, 1)) if new.shape[-1] == 4 else ()):
        new[:, :, col] = new[:, :, col] * (new_scale[axis] / orig_scale[axis])
        new[:, :, col] = torch.clamp(new[:, :, col], 0, new_scale[axis] - 1)
    return new.squeeze(0) if squeeze else new


def set_deterministic_seed(seed: int) -> None:
    """Seeds numpy, torch and the CUDA generator for evaluation (utils/misc.py:186-196; ``FPM_Pipeline.on_test_start`` does the
    same inline).  The cuDNN switch of the reference has no counterpart: cuDNN is not used."""
    print("SETTING SEED: ", seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
