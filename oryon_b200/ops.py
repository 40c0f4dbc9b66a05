"""Thin Python entry points to backbone building blocks of liboryon_b200.so (tests and kernel benchmarks).
Nothing here computes: every function forwards device pointers to the C ABI."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._torch_glue import as_device, device_of, ptr, stream_ptr

ACT = {"none": 0, "quickgelu": 1, "gelu": 2, "relu": 3}


def linear(A: Tensor, W: Tensor, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None, *, act: str = "none",
           alpha: float = 1.0, precision: int = 3) -> Tensor:
    """``residual + act(alpha * A @ W^T + bias)`` on the tcgen05 GEMM (``oryon_gemm_f32``).
    ``A [M,K]`` or ``[B,M,K]``, ``W [N,K]`` or ``[B,N,K]`` float32 -> float32 ``[.., M, N]``."""
    dev = device_of(A, W)
    A, W = as_device(A, dev, torch.float32), as_device(W, dev, torch.float32)
    batched = A.dim() == 3
    if batched != (W.dim() == 3):
        raise ValueError("linear: A and W must both be batched or both plain")
    batch = A.shape[0] if batched else 1
    M, K = A.shape[-2:]
    N = W.shape[-2]
    if W.shape[-1] != K or (batched and W.shape[0] != batch):
        raise ValueError(f"linear: shapes {tuple(A.shape)} x {tuple(W.shape)}")
    bias = None if bias is None else as_device(bias, dev, torch.float32)
    residual = None if residual is None else as_device(residual, dev, torch.float32)
    out = torch.empty(*A.shape[:-1], N, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().oryon_gemm_f32(_lib.handle(dev.index), ptr(A), ptr(W), ptr(bias), ptr(residual), ptr(out), M, N, K, batch,
                                          ACT[act], float(alpha), int(precision), stream_ptr(dev)))
    return out
