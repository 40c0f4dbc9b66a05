"""One CLIP-shaped GEMM a few times (ncu target: `ncu ... -k regex:gemm_tc2_kernel python tools/gemm_one.py`)."""
import faulthandler
import os
import sys

faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oryon_b200 import ops  # noqa: E402

PREC = int(os.environ.get("GEMM_ONE_PRECISION", "3"))
M = int(os.environ.get("GEMM_ONE_M", "18464"))
A = torch.randn(M, 1024, device="cuda")
W = torch.randn(3072, 1024, device="cuda") * 0.03
for _ in range(4):
    y = ops.linear(A, W, precision=PREC)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
