"""GPU parity: oryon_match_nn / nn_correspondences (through the C ABI) against the CPU oracle and the
golden fixtures produced by the unmodified reference.

Parity rule (integer/index work must be bit-exact; the distance is float32):
  * the chosen query index equals the reference's argmin wherever the float64 top-2 distance margin of
    that row exceeds MARGIN_TOL (the reference's own float32 evaluation order is not defined below that);
    on the remaining rows the chosen column must be distance-equivalent: |d(chosen) - min_dist| <= DIST_TOL;
  * min distance within DIST_TOL (float32 rounding of a D-term dot product of unit vectors).
"""
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from gpu_util import need_gpu
from oryon_b200 import _lib, synth
from oryon_b200.utils import pcd

pytestmark = pytest.mark.gpu

MARGIN_TOL = 2e-6
DIST_TOL = 1e-6


def _roi_feats(f, m):
    roi = torch.nonzero(m == 1)
    return roi, f[:, roi[:, 0], roi[:, 1]].T.float()


def _check_rows(idx, dist, ref_idx, ref_dist, f1, f2, margin=None):
    idx, dist = idx.cpu().long(), dist.cpu()
    assert idx.min() >= 0 and idx.max() < f2.shape[0]
    np.testing.assert_allclose(dist.numpy(), ref_dist.numpy(), atol=DIST_TOL, rtol=0)
    same = idx == ref_idx
    if margin is None:
        d64 = 0.5 * (1 - torch.nn.functional.normalize(f1.double(), dim=1) @ torch.nn.functional.normalize(f2.double(), dim=1).T)
        top2 = torch.topk(d64, k=min(2, d64.shape[1]), dim=1, largest=False)[0]
        margin = (top2[:, -1] - top2[:, 0]).float() if d64.shape[1] > 1 else torch.ones(d64.shape[0])
    clear = margin > MARGIN_TOL
    assert bool(same[clear].all()), f"{int((~same & clear).sum())} rows differ from the reference argmin with a clear margin"
    bad = torch.nonzero(~same).squeeze(1)
    if bad.numel():
        # distance-equivalent choice on (near-)ties
        a = torch.nn.functional.normalize(f1[bad].double(), dim=1)
        q = torch.nn.functional.normalize(f2[idx[bad]].double(), dim=1)
        d_chosen = 0.5 * (1 - (a * q).sum(1))
        assert bool(((d_chosen.float() - ref_dist[bad]).abs() <= DIST_TOL).all())
    return int((~same).sum())


@pytest.mark.parametrize("mode", [_lib.MATCH_TC_REFINED, _lib.MATCH_EXACT_FP32])
@pytest.mark.parametrize("case", list(synth.MATCH_CASES))
def test_match_rows_vs_reference_golden(golden_dir, case, mode):
    need_gpu()
    g = np.load(os.path.join(golden_dir, f"match_{case}.npz"))
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs(case)
    roi1, f1 = _roi_feats(fa, ma)
    roi2, f2 = _roi_feats(fq, mq)
    W1, W2 = fa.shape[2], fq.shape[2]
    pix1 = (roi1[:, 0] * W1 + roi1[:, 1]).int().cuda()
    pix2 = (roi2[:, 0] * W2 + roi2[:, 1]).int().cuda()
    # the library's own ROI enumeration equals torch.nonzero order
    p1, c1 = pcd.mask_to_roi(ma.cuda())
    assert int(c1[0]) == pix1.numel() and torch.equal(p1[0, :pix1.numel()], pix1)
    idx, dist = pcd.match_nn(fa[None].cuda(), fq[None].cuda(), pix1[None], pix2[None], [pix1.numel()], [pix2.numel()], mode=mode)
    torch.cuda.synchronize()
    n_diff = _check_rows(idx[0], dist[0], torch.from_numpy(g["nn_idx"]).long(), torch.from_numpy(g["min_dist"]), f1, f2,
                         torch.from_numpy(g["margin"]))
    print(case, "mode", mode, "rows", pix1.numel(), "near-tie index differences", n_diff, pcd.match_last_stats())


@pytest.mark.parametrize("case", list(synth.MATCH_CASES))
def test_nn_correspondences_vs_reference_golden(golden_dir, case):
    """The full reference function, RNG seeded as the reference seeds it: identical int64 rows."""
    need_gpu()
    g = np.load(os.path.join(golden_dir, f"match_{case}.npz"))
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs(case)
    torch.manual_seed(seed)
    corrs, dbg = pcd.nn_correspondences(fa.cuda(), fq.cuda(), ma.cuda(), mq.cuda(), th, max_corrs, sub, "cpu", return_debug=True)
    if bool(g["is_none"]):
        assert corrs is None
        return
    assert corrs is not None and corrs.dtype == torch.int64 and corrs.is_cuda and tuple(corrs.shape) == (max_corrs, 4)
    # oracle on the same inputs and the same draws
    torch.manual_seed(seed)
    ref, rdbg = oracle.nn_correspondences(fa, fq, ma, mq, th, max_corrs, sub, return_debug=True)
    assert np.array_equal(ref.numpy(), g["corrs"])
    same_rows = torch.equal(dbg["nn_idx"].cpu().long(), rdbg["nn_idx"]) and torch.equal(dbg["valid"].cpu(), rdbg["valid"])
    if same_rows:
        assert np.array_equal(corrs.cpu().numpy(), g["corrs"])
    else:  # a near-tie row picked a distance-equivalent neighbour: every returned row must still be a thresholded NN pair
        _check_corr_rows_are_nn_pairs(corrs.cpu(), fa, fq, ma, mq, th)


def _check_corr_rows_are_nn_pairs(corrs, fa, fq, ma, mq, th):
    """Every ``(y1,x1,y2,x2)`` row: both pixels inside their masks, the query pixel a nearest neighbour of the anchor pixel among
    the query ROI up to DIST_TOL (float64 evaluation), and its distance below the threshold."""
    roi2, f2 = _roi_feats(fq, mq)
    q = torch.nn.functional.normalize(f2.double(), dim=1)
    a = torch.nn.functional.normalize(fa[:, corrs[:, 0], corrs[:, 1]].T.double(), dim=1)
    d = 0.5 * (1 - a @ q.T)
    assert bool((ma[corrs[:, 0], corrs[:, 1]] == 1).all()) and bool((mq[corrs[:, 2], corrs[:, 3]] == 1).all())
    c = torch.nn.functional.normalize(fq[:, corrs[:, 2], corrs[:, 3]].T.double(), dim=1)
    d_chosen = 0.5 * (1 - (a * c).sum(1))
    assert bool((d_chosen - d.min(1).values <= DIST_TOL).all()) and bool((d_chosen < th + DIST_TOL).all())


def test_corrs_device_cuda_branch():
    """``corrs_device='cuda'`` (reference utils/pcd.py:195-197, utils/misc.py:242-254): the reference then evaluates the distances
    in FLOAT16 on the GPU and draws from the CUDA generator.  What parity means for this branch, and is asserted here:
      * draws: the two ``multinomial`` calls consume the default CUDA generator exactly as the reference's calls do (same
        weights vector ``ones(N, float64)`` on the device, same ``n``, same order), so with the distances below the sampled rows
        are reproduced by replaying those two draws;
      * distances: the library does NOT emulate the float16 branch -- it returns the float32-exact nearest neighbour (the
        'cpu' branch's arithmetic), which the float16 evaluation approximates: the float16 minimum differs from it by at most
        F16_TOL, and wherever a row's float32 top-2 margin exceeds 2 * F16_TOL both branches pick the same column.  (Below that
        margin the reference's own answer depends on float16 rounding and ATen's tie order.)"""
    need_gpu()
    F16_TOL = 2e-3
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs("m1_d32_64x64_subsample")
    fa_d, fq_d, ma_d, mq_d = fa.cuda(), fq.cuda(), ma.cuda(), mq.cuda()
    torch.manual_seed(seed)                      # seeds the CUDA generator too (utils/misc.py:194-195)
    corrs, dbg = pcd.nn_correspondences(fa_d, fq_d, ma_d, mq_d, th, max_corrs, sub, "cuda", return_debug=True)
    assert corrs is not None and corrs.is_cuda and corrs.dtype == torch.int64 and tuple(corrs.shape) == (max_corrs, 4)
    # the reference's float16 branch on the same device, with the same generator state
    torch.manual_seed(seed)
    roi1 = torch.nonzero(ma_d == 1)
    roi2 = torch.nonzero(mq_d == 1)
    assert roi1.shape[0] > sub
    idxs = pcd.torch_sample_select(roi1, sub)
    roi1 = roi1[idxs]
    f1 = fa_d[:, roi1[:, 0], roi1[:, 1]].T.to(torch.float16)
    f2 = fq_d[:, roi2[:, 0], roi2[:, 1]].T.to(torch.float16)
    d16 = 0.5 * (-1 * torch.nn.functional.cosine_similarity(f1.unsqueeze(1), f2.unsqueeze(0), dim=2) + 1)
    min16, arg16 = torch.amin(d16, dim=1), torch.argmin(d16, dim=1)
    # same source subsample (first draw identical), float32-exact distances close to the float16 ones
    W = fa.shape[2]
    assert torch.equal(dbg["pix1"].long(), roi1[:, 0] * W + roi1[:, 1])
    assert float((dbg["min_dist"] - min16.float()).abs().max()) <= F16_TOL
    d64 = 0.5 * (1 - torch.nn.functional.normalize(f1.double(), dim=1) @ torch.nn.functional.normalize(f2.double(), dim=1).T)
    top2 = torch.topk(d64, 2, dim=1, largest=False)[0]
    clear = (top2[:, 1] - top2[:, 0]) > 2 * F16_TOL
    assert int(clear.sum()) > 0 and torch.equal(dbg["nn_idx"].long()[clear], arg16[clear])
    # second draw replayed on the library's valid set: identical rows
    valid = torch.nonzero(dbg["min_dist"] < th).squeeze(1)
    assert torch.equal(valid, dbg["valid"])
    final = torch.cat((roi1[valid], roi2[dbg["nn_idx"].long()[valid]]), dim=1)
    sel = pcd.torch_sample_select(final, max_corrs)
    assert torch.equal(corrs, final[sel])
    _check_corr_rows_are_nn_pairs(corrs.cpu(), fa, fq, ma, mq, th)


@pytest.mark.parametrize("case", ["m0_d32_48x48", "m2_d128_40x40_full", "m5_d17_24x40_ragged", "m6_d32_96x96_refsize"])
def test_cta_pair_kernel_matches_reference_golden(golden_dir, case, monkeypatch):
    """The cta_group::2 variant of the tensor-core pass (ORYON_MATCH_PAIR=1; slower than the default at config 2, kept as a
    measured alternative) against the same reference goldens, plus a ragged dense batch against the oracle."""
    need_gpu()
    monkeypatch.setenv("ORYON_MATCH_PAIR", "1")
    g = np.load(os.path.join(golden_dir, f"match_{case}.npz"))
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs(case)
    roi1, f1 = _roi_feats(fa, ma)
    roi2, f2 = _roi_feats(fq, mq)
    pix1 = (roi1[:, 0] * fa.shape[2] + roi1[:, 1]).int().cuda()
    pix2 = (roi2[:, 0] * fq.shape[2] + roi2[:, 1]).int().cuda()
    idx, dist = pcd.match_nn(fa[None].cuda(), fq[None].cuda(), pix1[None], pix2[None], [pix1.numel()], [pix2.numel()])
    torch.cuda.synchronize()
    _check_rows(idx[0], dist[0], torch.from_numpy(g["nn_idx"]).long(), torch.from_numpy(g["min_dist"]), f1, f2, torch.from_numpy(g["margin"]))
    fa, fq, perm = synth.permuted_feature_batch(77, 3, 64, 30, 34, noise=0.3)
    idx, dist = pcd.match_nn(fa.cuda(), fq.cuda())
    torch.cuda.synchronize()
    for b in range(3):
        a, q = fa[b].reshape(64, -1).T.contiguous(), fq[b].reshape(64, -1).T.contiguous()
        rd, ri = oracle.match_rows(a, q)
        _check_rows(idx[b], dist[b], ri, rd, a, q)


def test_host_tensors_are_accepted_and_result_returns_to_host():
    need_gpu()
    fa, fq, ma, mq, th, max_corrs, sub, seed = synth.match_inputs("m0_d32_48x48")
    torch.manual_seed(seed)
    c_host = pcd.nn_correspondences(fa, fq, ma, mq, th, max_corrs, sub, "cpu")
    torch.manual_seed(seed)
    c_dev = pcd.nn_correspondences(fa.cuda(), fq.cuda(), ma.cuda(), mq.cuda(), th, max_corrs, sub, "cpu")
    assert c_host.device.type == "cpu" and torch.equal(c_host, c_dev.cpu())


@pytest.mark.parametrize("D,h,w,B", [(128, 40, 40, 3), (256, 24, 24, 2), (17, 20, 30, 2), (1, 16, 16, 1), (64, 33, 17, 2), (96, 31, 9, 1),
                                     (200, 16, 16, 1), (32, 24, 26, 2), (64, 30, 34, 2)])
def test_dense_batched_vs_oracle(D, h, w, B):
    need_gpu()
    fa, fq, perm = synth.permuted_feature_batch(7 + D, B, D, h, w, noise=0.3)
    idx, dist = pcd.match_nn(fa.cuda(), fq.cuda())
    idx_e, dist_e = pcd.match_nn(fa.cuda(), fq.cuda(), mode=_lib.MATCH_EXACT_FP32)
    torch.cuda.synchronize()
    for b in range(B):
        f1, f2 = fa[b].reshape(D, -1).T.contiguous(), fq[b].reshape(D, -1).T.contiguous()
        rd, ri = oracle.match_rows(f1, f2)
        _check_rows(idx[b], dist[b], ri, rd, f1, f2)
        _check_rows(idx_e[b], dist_e[b], ri, rd, f1, f2)


def test_ragged_batch_and_empty_lists():
    need_gpu()
    B, D, h, w = 4, 32, 24, 24
    fa, fq, _ = synth.permuted_feature_batch(11, B, D, h, w, noise=0.2)
    g = torch.Generator().manual_seed(5)
    n_a, n_q = [300, 0, 129, 576], [200, 50, 0, 576]
    roi_a = torch.stack([torch.randperm(h * w, generator=g) for _ in range(B)]).int()
    roi_q = torch.stack([torch.randperm(h * w, generator=g) for _ in range(B)]).int()
    idx, dist = pcd.match_nn(fa.cuda(), fq.cuda(), roi_a.cuda(), roi_q.cuda(), n_a, n_q)
    torch.cuda.synchronize()
    idx, dist = idx.cpu(), dist.cpu()
    for b in range(B):
        assert bool((idx[b, n_a[b]:] == -1).all()) and bool(torch.isinf(dist[b, n_a[b]:]).all())
        if n_a[b] == 0:
            continue
        if n_q[b] == 0:
            assert bool((idx[b, :n_a[b]] == -1).all())
            continue
        f1 = fa[b].reshape(D, -1)[:, roi_a[b, :n_a[b]].long()].T.contiguous()
        f2 = fq[b].reshape(D, -1)[:, roi_q[b, :n_q[b]].long()].T.contiguous()
        rd, ri = oracle.match_rows(f1, f2)
        _check_rows(idx[b, :n_a[b]], dist[b, :n_a[b]], ri, rd, f1, f2)


def test_exact_ties_pick_lowest_index_and_zero_rows_never_match():
    need_gpu()
    D, n = 32, 700
    g = torch.Generator().manual_seed(3)
    q = torch.randn(n, D, generator=g)
    q[400:] = q[:300]                      # every column 400+i duplicates column i (exact ties)
    a = q[torch.randint(0, 300, (512,), generator=g)].clone()
    a[7] = 0.0                             # zero feature row: cos = 0 -> dist 0.5 (never passes th=0.25)
    fa = a.T.reshape(1, D, 512, 1).contiguous()
    fq = q.T.reshape(1, D, n, 1).contiguous()
    for mode in (_lib.MATCH_TC_REFINED, _lib.MATCH_EXACT_FP32):
        idx, dist = pcd.match_nn(fa.cuda(), fq.cuda(), mode=mode)
        idx, dist = idx[0].cpu().long(), dist[0].cpu()
        rd, ri = oracle.match_rows(a, q)
        keep = torch.arange(512) != 7
        assert torch.equal(idx[keep], ri[keep]) and bool((idx[keep] < 300).all())
        assert abs(float(dist[7]) - 0.5) < 1e-6
        np.testing.assert_allclose(dist.numpy(), rd.numpy(), atol=DIST_TOL)


def test_overflowing_candidate_lists_fall_back_to_exact_rows():
    """Hundreds of near-identical query columns exceed the per-row candidate capacity: the row must be
    finished by the exact kernel and still return the first best column."""
    need_gpu()
    D, n = 32, 2048
    g = torch.Generator().manual_seed(9)
    base = torch.randn(1, D, generator=g)
    q = base + 1e-4 * torch.randn(n, D, generator=g)
    a = base + 1e-4 * torch.randn(256, D, generator=g)
    idx, dist = pcd.match_nn(a.T.reshape(1, D, 256, 1).contiguous().cuda(), q.T.reshape(1, D, n, 1).contiguous().cuda())
    torch.cuda.synchronize()
    stats = pcd.match_last_stats()
    assert stats["rows_overflowed"] > 0
    rd, ri = oracle.match_rows(a, q)
    _check_rows(idx[0], dist[0], ri, rd, a, q)


def test_config2_full_size_recovers_planted_permutation():
    """BASELINE config 2 at full size (B=32, D=128, 120x160): the oracle cannot finish this in seconds, so
    parity is checked through the planted permutation (every anchor pixel has exactly one strong match)
    and by re-evaluating the reported distance of the reported neighbour."""
    need_gpu()
    B, D, h, w = 32, 128, 120, 160
    fa, fq, perm = synth.permuted_feature_batch(0, B, D, h, w, noise=0.1, device="cuda")
    idx, dist = pcd.match_nn(fa, fq)
    torch.cuda.synchronize()
    assert torch.equal(idx.long(), perm)
    a = torch.nn.functional.normalize(fa.view(B, D, -1), dim=1)
    q = torch.nn.functional.normalize(torch.gather(fq.view(B, D, -1), 2, perm[:, None, :].expand(B, D, -1)), dim=1)
    d = 0.5 * (1 - (a * q).sum(1))
    assert float((d - dist).abs().max()) <= 2e-6
    assert pcd.match_last_stats()["rows_overflowed"] == 0


def test_config5_full_size_properties():
    """BASELINE config 5 (D=256, 240x320 -> 76 800 positions per image; 2 pairs = 6 TFLOP) through size-independent
    properties: the planted permutation is recovered, the reported distance equals a float32 re-evaluation of the reported
    neighbour, and matching a map against itself is the identity at distance ~0 (idempotence)."""
    need_gpu()
    B, D, h, w = 2, 256, 240, 320
    fa, fq, perm = synth.permuted_feature_batch(5, B, D, h, w, noise=0.1, device="cuda")
    idx, dist = pcd.match_nn(fa, fq)
    torch.cuda.synchronize()
    assert torch.equal(idx.long(), perm)
    a = torch.nn.functional.normalize(fa.view(B, D, -1), dim=1)
    q = torch.nn.functional.normalize(torch.gather(fq.view(B, D, -1), 2, perm[:, None, :].expand(B, D, -1)), dim=1)
    assert float((0.5 * (1 - (a * q).sum(1)) - dist).abs().max()) <= 2e-6
    del a, q, fq
    idx2, dist2 = pcd.match_nn(fa, fa)
    torch.cuda.synchronize()
    n = h * w
    assert torch.equal(idx2.long(), torch.arange(n, device="cuda")[None].expand(B, n))
    assert float(dist2.abs().max()) <= 2e-6
