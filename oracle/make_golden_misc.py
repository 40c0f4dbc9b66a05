"""Generate ``tests/golden/misc_0.npz``: outputs of the reference's ``utils.misc.rescale_coords`` (:93-122) on seeded inputs
(float and int64 coordinates, 2 and 4 columns, single and batched, values that need the clamp).

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_misc.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

ref_shims.install()
from utils.misc import rescale_coords  # noqa: E402  (reference)

from oryon_b200 import synth  # noqa: E402


def main():
    out = {}
    for name, (coords, a, b) in synth.rescale_cases(0).items():
        out[name] = rescale_coords(coords, a, b).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "misc_0.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
