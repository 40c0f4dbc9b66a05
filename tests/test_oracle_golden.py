"""CPU: the oracle restatement must reproduce the outputs the UNMODIFIED reference produced
(tests/golden/*.npz, written by oracle/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from oryon_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("case", list(synth.MATCH_CASES))
def test_match_oracle_equals_reference(golden_dir, case):
    g = _load(golden_dir, f"match_{case}.npz")
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs(case)
    sums = [synth.tensor_checksum(t) for t in (fa, fq, ma, mq)]
    assert sums == list(g["in_sum"]), "synthetic inputs drifted from the ones the fixture was made with"

    # deterministic part: row-wise min distance and argmin, bit-exact (same torch ops as the reference)
    roi1, roi2 = torch.nonzero(ma == 1), torch.nonzero(mq == 1)
    assert roi1.shape[0] == int(g["n1"]) and roi2.shape[0] == int(g["n2"])
    f1 = fa[:, roi1[:, 0], roi1[:, 1]].T.float()
    f2 = fq[:, roi2[:, 0], roi2[:, 1]].T.float()
    min_dist, nn_idx = oracle.match_rows(f1, f2, row_chunk=97)  # ragged chunk on purpose
    assert np.array_equal(nn_idx.numpy(), g["nn_idx"].astype(np.int64))
    assert np.array_equal(min_dist.numpy(), g["min_dist"])
    # un-chunked distance matrix is identical too
    full = oracle.inv_norm_cosine(f1, f2)
    assert torch.equal(torch.amin(full, 1), min_dist)

    # full function incl. the RNG draws on the global CPU generator, seeded like the reference
    torch.manual_seed(seed)
    corrs = oracle.nn_correspondences(fa, fq, ma, mq, th, max_corrs, sub)
    if bool(g["is_none"]):
        assert corrs is None
    else:
        assert corrs.dtype == torch.int64
        assert np.array_equal(corrs.numpy(), g["corrs"])


@pytest.mark.parametrize("seed", list(synth.LIFT_CASES))
def test_lift_oracle_equals_reference(golden_dir, seed):
    g = _load(golden_dir, f"lift_{seed}.npz")
    corrs, depth_a, depth_q, K, fm, raw = synth.lift_inputs(seed)
    assert [synth.tensor_checksum(t) for t in (corrs, depth_a, depth_q)] == list(g["in_sum"])
    pa, pq = oracle.corrs_to_pcds(corrs, depth_a, depth_q, K, K, fm, raw, raw)
    assert pa.dtype == torch.float32
    assert np.array_equal(pa.numpy(), g["pcd_a"])
    assert np.array_equal(pq.numpy(), g["pcd_q"])
    ca = oracle.scale_coords(corrs[:, :2], fm, raw)
    cq = oracle.scale_coords(corrs[:, 2:], fm, raw)
    ok = oracle.get_valid_coords(ca, raw) & oracle.get_valid_coords(cq, raw)
    assert np.array_equal(ca[ok].long().numpy(), g["ca"].astype(np.int64))


@pytest.mark.parametrize("seed", list(synth.POINTDSC_CASES))
def test_pointdsc_oracle_equals_reference(golden_dir, seed):
    g = _load(golden_dir, f"pointdsc_{seed}.npz")
    n, out_frac = synth.POINTDSC_CASES[seed]
    sd = synth.pointdsc_state_dict(seed)
    data = synth.rigid_correspondences(seed, n=n, outlier_frac=out_frac)
    assert [synth.tensor_checksum(t) for t in (data["src"], data["tgt"], sd["encoder.layer0.weight"])] == list(g["in_sum"])
    torch.set_num_threads(8)
    final, dbg = oracle.pointdsc_pose(sd, synth.POINTDSC_DEFAULT_CFG, data["src"], data["tgt"], return_debug=True)
    # float32 network, different (but equivalent) op grouping than nn.Module: tight tolerance, not bitwise
    np.testing.assert_allclose(dbg["sc"][:8, :8].numpy(), g["sc_sample"], atol=1e-6)
    np.testing.assert_allclose(dbg["feat"][:, :8].numpy(), g["feat_sample"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(dbg["conf"].numpy(), g["conf"], rtol=2e-4, atol=2e-4)
    assert np.array_equal(dbg["seeds"].numpy(), g["seeds"])
    np.testing.assert_allclose(final.numpy(), g["final_trans"], atol=1e-5)


def test_mask_postproc_matches_torch_semantics():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(2, 1, 12, 12, generator=g)
    gt = (torch.rand(2, 12, 12, generator=g) > 0.5).float()
    m = oracle.predicted_mask(logits)
    assert m.dtype == torch.int64 and torch.equal(m, (logits.squeeze(1) > 0).long())
    iou = oracle.mask_iou(gt, m)
    ref = torch.stack([((gt[i] > 0) & (m[i] > 0)).sum() / ((gt[i] > 0) | (m[i] > 0)).sum() for i in range(2)])
    assert torch.equal(iou, ref)
