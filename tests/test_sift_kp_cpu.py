"""Key-point form of the matcher (the reference's SIFT baselines, scripts/evaluation/sift_nocs.py:25-45 / sift_toyl.py:25-51):
the oracle restatement against the outputs of the reference's own two functions on real SIFT descriptor sets
(``oracle/make_golden_sift.py`` -> ``tests/golden/sift_kp_*.npz``)."""
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from oryon_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(case):
    g = np.load(os.path.join(GOLD, f"sift_kp_{case}.npz"))
    f1, f2 = torch.tensor(g["f1"].astype(np.float32)), torch.tensor(g["f2"].astype(np.float32))
    return g, f1, f2, torch.tensor(g["k1"]), torch.tensor(g["k2"])


@pytest.mark.parametrize("case", list(synth.SIFT_CASES))
def test_kp_oracle_equals_reference(case):
    g, f1, f2, k1, k2 = load_case(case)
    th, seed = float(g["threshold"]), int(g["seed"])
    min_dist, nn_idx = oracle.match_rows(f1, f2)
    assert np.array_equal(nn_idx.numpy(), g["nn_idx"]) and np.array_equal(min_dist.numpy(), g["min_dist"])
    torch.manual_seed(seed)
    toyl = oracle.nn_correspondences_kp(f1, f2, k1, k2, th, 500, max_source=1000, keep_empty=True)
    assert toyl.dtype == torch.int16 and np.array_equal(toyl.numpy(), g["corrs_toyl"])
    torch.manual_seed(seed)
    if bool(g["nocs_raises"]):
        with pytest.raises(RuntimeError):
            oracle.nn_correspondences_kp(f1, f2, k1, k2, th, 500)
    else:
        assert np.array_equal(oracle.nn_correspondences_kp(f1, f2, k1, k2, th, 500).numpy(), g["corrs_nocs"])


def _oracle_match_nn(fa, fq, roi_a, roi_q, n_a, n_q):
    """Stand-in for ``pcd.match_nn`` (the library call) with its contract: positions in the query list, -1 / inf past n_a."""
    n1, n2 = n_a[0], n_q[0]
    f1, f2 = fa[0][:, roi_a[0, :n1].long()].T.contiguous(), fq[0][:, roi_q[0, :n2].long()].T.contiguous()
    d, i = oracle.match_rows(f1, f2)
    idx = torch.full((1, roi_a.shape[1]), -1, dtype=torch.int32)
    dist = torch.full((1, roi_a.shape[1]), float("inf"))
    idx[0, :n1], dist[0, :n1] = i.int(), d
    return idx, dist


@pytest.mark.parametrize("variant", ["nocs", "toyl"])
@pytest.mark.parametrize("case", list(synth.SIFT_CASES))
def test_kp_host_logic_with_oracle_matcher(monkeypatch, case, variant):
    """Everything of ``pcd.nn_correspondences_kp`` around the library call (set packing, position lists, subsample / final
    draws, key-point rows, empty-set behaviour) reproduces the reference's rows when the call is answered by the oracle."""
    from oryon_b200.utils import pcd
    monkeypatch.setattr(pcd, "match_nn", _oracle_match_nn)
    monkeypatch.setattr(pcd, "device_of", lambda *t: torch.device("cpu"))
    g, f1, f2, k1, k2 = load_case(case)
    kw = dict(max_source=1000, keep_empty=True) if variant == "toyl" else {}
    torch.manual_seed(int(g["seed"]))
    if variant == "nocs" and bool(g["nocs_raises"]):
        with pytest.raises(RuntimeError):
            pcd.nn_correspondences_kp(f1, f2, k1, k2, float(g["threshold"]), 500)
        return
    got = pcd.nn_correspondences_kp(f1, f2, k1, k2, float(g["threshold"]), 500, **kw)
    assert got.dtype == torch.int16 and np.array_equal(got.numpy(), g[f"corrs_{variant}"])
