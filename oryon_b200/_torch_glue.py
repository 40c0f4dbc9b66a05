"""Torch-side plumbing shared by the interface mirrors: device pointers, the current stream and
argument normalisation.  No arithmetic happens here."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import Tensor

from . import _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.OryonError("oryon_b200 needs a CUDA device (sm_100); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device(t: Tensor, device: torch.device, dtype: Optional[torch.dtype] = None) -> Tensor:
    """Contiguous copy/view of ``t`` on ``device`` (H2D copy if it lives on the host)."""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def ptr(t: Optional[Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def device_of(*tensors: Tensor) -> torch.device:
    """The CUDA device the call runs on: that of the first CUDA tensor argument, else the current one."""
    for t in tensors:
        if isinstance(t, Tensor) and t.is_cuda:
            return t.device
    return require_cuda()
