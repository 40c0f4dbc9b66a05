"""CPU: ``oryon_b200.utils.misc`` against the reference's ``utils.misc`` helpers (``oracle/make_golden_misc.py`` -> tests/golden/misc_0.npz)."""
import os

import numpy as np
import torch

from oryon_b200 import synth
from oryon_b200.utils import misc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rescale_coords_equals_reference():
    g = np.load(os.path.join(GOLD, "misc_0.npz"))
    for name, (coords, a, b) in synth.rescale_cases(0).items():
        before = coords.clone()
        out = misc.rescale_coords(coords, a, b)
        assert out.dtype == coords.dtype and torch.equal(coords, before), "returns a copy"
        assert np.array_equal(out.numpy(), g[name]), name


def test_set_deterministic_seed_seeds_numpy_and_torch(capsys):
    misc.set_deterministic_seed(7)
    a = (np.random.randint(0, 1000), torch.rand(3))
    misc.set_deterministic_seed(7)
    b = (np.random.randint(0, 1000), torch.rand(3))
    assert a[0] == b[0] and torch.equal(a[1], b[1]) and os.environ["PYTHONHASHSEED"] == "7"
    assert misc.torch_sample_select(torch.zeros(10, 2), 4).shape == (4,)
