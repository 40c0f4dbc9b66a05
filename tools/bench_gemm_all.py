"""Kernel-only timing of gemm_tc on every GEMM shape of one network pass (N = 32 images), precision 3 and 1."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oryon_b200 import _lib, ops  # noqa: E402

N = 32
SHAPES = [  # name, count per pass, batch, M, N, K
    ("clip_patch", 1, 1, N * 576, 1024, 588), ("clip_qkv", 24, 1, N * 577, 3072, 1024), ("clip_qk", 24, N * 16, 577, 577, 64),
    ("clip_pv", 24, N * 16, 577, 64, 577), ("clip_out", 24, 1, N * 577, 1024, 1024), ("clip_fc", 24, 1, N * 577, 4096, 1024),
    ("clip_proj", 24, 1, N * 577, 1024, 4096),
    ("swin_patch", 1, 1, N * 9216, 128, 48), ("swin1_qkv", 2, 1, N * 9604, 384, 128), ("swin1_proj", 2, 1, N * 9604, 128, 128),
    ("swin1_fc1", 2, 1, N * 9216, 512, 128), ("swin1_fc2", 2, 1, N * 9216, 128, 512), ("swin_merge1", 1, 1, N * 2304, 256, 512),
    ("swin2_qkv", 2, 1, N * 2401, 768, 256), ("swin2_proj", 2, 1, N * 2401, 256, 256), ("swin2_fc1", 2, 1, N * 2304, 1024, 256),
    ("swin2_fc2", 2, 1, N * 2304, 256, 1024), ("swin_merge2", 1, 1, N * 576, 512, 1024),
    ("f_clipconv", 1, 1, N * 576, 768, 1024), ("f_corr", 2, N // 2, 576, 80, 768), ("f_conv1", 1, 1, N * 576, 128, 3920),
    ("f_guid", 1, 1, N * 576, 128, 4608), ("f_qkv", 4, 1, N * 576, 384, 256), ("f_proj", 4, 1, N * 576, 128, 128),
    ("f_fc1", 4, 1, N * 576, 512, 128), ("f_fc2", 4, 1, N * 576, 128, 512),
    ("d_g0", 1, 1, N * 2304, 32, 2304), ("d_g1", 1, 1, N * 9216, 16, 1152),
    ("d1_up", 1, 1, N * 576, 384, 128), ("d1_c0", 1, 1, N * 2304, 64, 1152), ("d1_c1", 1, 1, N * 2304, 64, 576),
    ("d2_up", 1, 1, N * 2304, 192, 64), ("d2_c0", 1, 1, N * 9216, 32, 576), ("d2_c1", 1, 1, N * 9216, 32, 288),
    ("d3_up", 1, 1, N * 9216, 128, 32), ("d3_c0", 1, 1, N * 36864, 32, 288), ("d3_c1", 1, 1, N * 36864, 32, 288),
]


def main():
    torch.cuda.set_device(0)
    res, tot = {}, {1: 0.0, 3: 0.0}
    for name, count, batch, M, Nn, K in SHAPES:
        A = torch.randn((batch, M, K) if batch > 1 else (M, K), device="cuda")
        W = torch.randn((batch, Nn, K) if batch > 1 else (Nn, K), device="cuda") * 0.03
        row = {}
        for prec in (3, 1):
            for _ in range(2):
                ops.linear(A, W, precision=prec)
            torch.cuda.synchronize()
            _lib.profile_enable(0, True)
            _lib.profile_read(0)
            for _ in range(3):
                ops.linear(A, W, precision=prec)
            torch.cuda.synchronize()
            p = _lib.profile_read(0)
            _lib.profile_enable(0, False)
            ms = p["gemm_tc"][0] / p["gemm_tc"][1]
            row[f"p{prec}_ms"] = round(ms, 4)
            row[f"p{prec}_tflops"] = round(2.0 * batch * M * Nn * K / ms / 1e9, 1)
            row[f"p{prec}_total_ms"] = round(ms * count, 3)
            tot[prec] += ms * count
        res[name] = row
        del A, W
    res["_total_ms"] = {f"p{k}": round(v, 2) for k, v in tot.items()}
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()
