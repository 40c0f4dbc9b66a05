"""GPU parity (bit-exact): coordinate scaling + bounds + truncation + lifting against the golden
fixtures written by the unmodified reference (pipeline.py:447-460, utils/pcd.py:35-81)."""
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from gpu_util import need_gpu
from oryon_b200 import synth
from oryon_b200.utils import pcd

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(synth.LIFT_CASES))
def test_corrs_to_pcd_bit_exact(golden_dir, seed):
    need_gpu()
    g = np.load(os.path.join(golden_dir, f"lift_{seed}.npz"))
    corrs, depth_a, depth_q, K, fm, raw = synth.lift_inputs(seed)
    pa, pq = pcd.corrs_to_pcd(corrs.cuda(), depth_a.cuda(), depth_q.cuda(), K, K, fm, raw, raw)
    assert pa.dtype == torch.float32 and pa.shape == g["pcd_a"].shape
    assert np.array_equal(pa.cpu().numpy(), g["pcd_a"])
    assert np.array_equal(pq.cpu().numpy(), g["pcd_q"])


@pytest.mark.parametrize("dtype", [torch.int32, torch.float32, torch.int16])
def test_lift_pcd_bit_exact_vs_oracle(dtype):
    need_gpu()
    corrs, depth_a, _, K, fm, raw = synth.lift_inputs(200)
    gen = torch.Generator().manual_seed(1)
    xs = torch.randint(0, raw[1], (777,), generator=gen)
    ys = torch.randint(0, raw[0], (777,), generator=gen)
    d = depth_a.to(dtype)
    ref = oracle.lift_pcd(d.unsqueeze(-1), K, (xs, ys)).float()
    got = pcd.lift_pcd(d.unsqueeze(-1).cuda(), K, (xs.cuda(), ys.cuda()))
    assert np.array_equal(got.cpu().numpy(), ref.numpy())


def test_lift_full_image_order():
    need_gpu()
    _, depth_a, _, K, _, raw = synth.lift_inputs(201)
    got = pcd.lift_pcd(depth_a.unsqueeze(-1).cuda(), K)
    ys, xs = torch.meshgrid(torch.arange(raw[0]), torch.arange(raw[1]), indexing="ij")
    ref = oracle.lift_pcd(depth_a.unsqueeze(-1), K, (xs.flatten(), ys.flatten())).float()
    assert np.array_equal(got.cpu().numpy(), ref.numpy())


def test_no_valid_rows():
    need_gpu()
    corrs = torch.full((10, 4), 10_000, dtype=torch.int64)
    depth = torch.ones(48, 64, dtype=torch.int32)
    K = torch.tensor(synth.NOCS_INTRINSICS, dtype=torch.float64)
    pa, pq = pcd.corrs_to_pcd(corrs.cuda(), depth.cuda(), depth.cuda(), K, K, (192, 192), (48, 64), (48, 64))
    assert pa.shape == (0, 3) and pq.shape == (0, 3)


@pytest.mark.parametrize("dtype", [torch.int32, torch.int16, torch.float32])
def test_select_lift_batched_bit_exact_vs_oracle(dtype):
    """oryon_select_lift (B pairs, one launch) == per-pair row selection (utils/pcd.py:207-212) followed by the
    oracle's corrs_to_pcds (pipeline.py:447-460), including a pair without correspondences and out-of-bounds rows."""
    need_gpu()
    B, n, W, cap_a, cap_q = 5, 500, 192, 3000, 4000
    raw = (120, 200)
    FM = (150, W)                                   # nominal map height < 192: rows with y >= 150 scale out of bounds
    gen = torch.Generator().manual_seed(3)
    roi_a = torch.stack([torch.randperm(192 * W, generator=gen)[:cap_a].sort().values for _ in range(B)]).int()
    roi_q = torch.stack([torch.randperm(192 * W, generator=gen)[:cap_q].sort().values for _ in range(B)]).int()
    nn_idx = torch.randint(0, cap_q, (B, cap_a), generator=gen).int()
    rows = torch.randint(0, cap_a, (B, n), generator=gen).int()
    rows[2] = -1
    depth_a = torch.randint(400, 1600, (B, *raw), generator=gen).to(dtype)
    depth_q = torch.randint(400, 1600, (B, *raw), generator=gen).to(dtype)
    K = torch.tensor(synth.NOCS_INTRINSICS, dtype=torch.float64).reshape(1, 3, 3).repeat(B, 1, 1)
    K[:, 0, 0] += torch.arange(B, dtype=torch.float64) * 3.25          # per-pair intrinsics
    Kq = K.clone()
    Kq[:, 1, 2] -= 7.5
    corrs, pa, pq, nv = pcd.select_lift_batched(rows.cuda(), roi_a.cuda(), roi_q.cuda(), nn_idx.cuda(), depth_a.cuda(), depth_q.cuda(),
                                                K, Kq, FM)
    nv = nv.cpu().tolist()
    assert nv[2] == -1
    for b in range(B):
        if b == 2:
            continue
        r = rows[b].long()
        p1, p2 = roi_a[b, r].long(), roi_q[b, nn_idx[b, r].long()].long()
        ref_c = torch.stack((p1 // W, p1 % W, p2 // W, p2 % W), dim=1)
        assert torch.equal(corrs[b].cpu(), ref_c)
        ra, rq = oracle.corrs_to_pcds(ref_c, depth_a[b], depth_q[b], K[b], Kq[b], FM, raw, raw)
        assert 0 < ra.shape[0] < n and nv[b] == ra.shape[0]
        assert np.array_equal(pa[b, :nv[b]].cpu().numpy(), ra.float().numpy())
        assert np.array_equal(pq[b, :nv[b]].cpu().numpy(), rq.float().numpy())
