"""Mirror of the reference ``utils/coordinates.py`` helpers that sit on the hot path (:5-13, :36-47).

``FPM_Pipeline.get_pose`` does not call these: it uses the fused library kernel
(``oryon_b200.utils.pcd.corrs_to_pcd`` -> ``oryon_corrs_to_pcd``), which performs the same float32
multiply, bounds test and truncation on the GPU.  They are kept, with the reference's signatures, as
shape/dtype plumbing for callers that index with them (e.g. scripts/evaluation/sift_*.py).
"""
from typing import Tuple, Union

import torch
from torch import Tensor


def scale_coords(coords: Tensor, source_scale: Union[Tensor, Tuple], target_scale: Union[Tensor, Tuple]) -> Tensor:
    """All measures are Y,X.  Returns a float32 copy (reference utils/coordinates.py:5-13)."""
    new_coords = coords.clone().to(torch.float32)
    new_coords[:, 0] = new_coords[:, 0] * (target_scale[0] / source_scale[0])
    new_coords[:, 1] = new_coords[:, 1] * (target_scale[1] / source_scale[1])
    return new_coords


def get_valid_coords(coords: Tensor, bounds: Union[Tensor, Tuple]) -> Tensor:
    """Boolean mask of coordinates inside ``bounds`` (Y,X) (reference utils/coordinates.py:36-47)."""
    ys, xs = coords[:, 0], coords[:, 1]
    return torch.logical_and(torch.logical_and(xs >= 0, xs < bounds[1]), torch.logical_and(ys >= 0, ys < bounds[0]))
