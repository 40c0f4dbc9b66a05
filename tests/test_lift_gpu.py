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
