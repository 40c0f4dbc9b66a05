"""Kernel-only timing of gemm_tc on the backbone's shapes (CUDA events recorded by the library around the launch)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oryon_b200 import _lib, ops  # noqa: E402

SHAPES = [  # name, M, N, K, act, residual
    ("clip_qkv", 18464, 3072, 1024, "none", False),
    ("clip_out", 18464, 1024, 1024, "none", True),
    ("clip_fc", 18464, 4096, 1024, "quickgelu", False),
    ("clip_proj", 18464, 1024, 4096, "none", True),
    ("dec_conv192", 1179648, 32, 288, "none", False),
    ("swin_qkv", 307328, 384, 128, "none", False),
]


def main():
    torch.cuda.set_device(0)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    res = {}
    for name, M, N, K, act, use_res in SHAPES:
        if only and name != only:
            continue
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda") * 0.03
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if use_res else None
        for prec in (1, 3):
            for _ in range(2):
                ops.linear(A, W, b, r, act=act, precision=prec)
            torch.cuda.synchronize()
            _lib.profile_enable(0, True)
            for _ in range(5):
                ops.linear(A, W, b, r, act=act, precision=prec)
            torch.cuda.synchronize()
            p = _lib.profile_read(0)
            _lib.profile_enable(0, False)
            ms = p["gemm_tc"][0] / p["gemm_tc"][1]
            res[f"{name}_p{prec}"] = dict(ms=round(ms, 4), tflops=round(2.0 * M * N * K / ms / 1e9, 1))
        del A, W, b, r
    print(json.dumps(res))


if __name__ == "__main__":
    main()
