"""Progress markers around the first library calls (diagnosis of an application crash under ncu on some boxes)."""
import faulthandler
import os
import sys

faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def say(*a):
    print(*a, flush=True)
    sys.stderr.flush()


import torch  # noqa: E402
say("torch imported", torch.__version__)
from oryon_b200 import _lib, ops  # noqa: E402
M = int(os.environ.get("GEMM_ONE_M", "36928"))
A = torch.randn(M, 1024, device="cuda")
W = torch.randn(3072, 1024, device="cuda") * 0.03
torch.cuda.synchronize()
say("inputs ready")
lib = _lib.load()
say("library loaded, abi", lib.oryon_abi_version())
h = _lib.handle(0)
say("handle ok")
for prec in (1, 3, 2):
    y = ops.linear(A, W, precision=prec)
    torch.cuda.synchronize()
    say("linear precision", prec, "ok", float(y[0, 0]))
say("done")
