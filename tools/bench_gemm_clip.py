"""Kernel-only timing of the four CLIP vision linear layers of one 32-pair pass (M = 64 x 577 rows) on the CTA-pair GEMM at precision
1 / 2 / 3, with the full operand ring and -- ORYON_GEMM_PAIR_STAGES=2 -- with a two-stage ring (half the operand bytes in flight):
separates "L2 -> SM bandwidth" from "latency x bytes in flight" as the bound of the feed.  Prints ms, TFLOP/s (algorithmic and
tensor-pipe, fp16-equivalent) and the operand bytes the pairs pull from L2 per second."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oryon_b200 import _lib, ops  # noqa: E402

M = 64 * 577
SHAPES = [("qkv", 3072, 1024), ("out", 1024, 1024), ("fc", 4096, 1024), ("proj", 1024, 4096)]
UNITS = {1: 1, 2: 2, 3: 3}


def timed(A, W, prec, reps=5):
    for _ in range(2):
        ops.linear(A, W, precision=prec)
    torch.cuda.synchronize()
    _lib.profile_enable(0, True)
    _lib.profile_read(0)
    for _ in range(reps):
        ops.linear(A, W, precision=prec)
    torch.cuda.synchronize()
    p = _lib.profile_read(0)
    _lib.profile_enable(0, False)
    return p["gemm_tc"][0] / p["gemm_tc"][1]


def main():
    torch.cuda.set_device(0)
    out = []
    for name, N, K in SHAPES:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda") * 0.03
        for prec in (1, 2, 3):
            for stages in ("full", "2"):
                if stages == "2" and prec == 1:
                    continue
                if stages == "2":
                    os.environ["ORYON_GEMM_PAIR_STAGES"] = "2"
                else:
                    os.environ.pop("ORYON_GEMM_PAIR_STAGES", None)
                ms = timed(A, W, prec)
                flops = 2.0 * M * N * K
                # every 256 x 256 x 64 unit pulls 512 rows x 64 K x (2 B hi [+ 2 B lo]) from L2
                units = -(-M // 256) * (N // 256) * (K // 64)
                l2_bytes = units * 512 * 64 * (2 if prec == 1 else 4)
                out.append(dict(shape=name, N=N, K=K, precision=prec, ring=stages, ms=round(ms, 4), algorithmic_tflops=round(flops / ms / 1e9, 1),
                                tensor_tflops_f16eq=round(UNITS[prec] * flops / ms / 1e9, 1), l2_to_sm_tbs=round(l2_bytes / ms / 1e9, 2)))
                print(json.dumps(out[-1]), flush=True)
        del A, W
    os.environ.pop("ORYON_GEMM_PAIR_STAGES", None)


if __name__ == "__main__":
    main()
