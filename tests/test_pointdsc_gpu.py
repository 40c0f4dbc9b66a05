"""GPU parity: PointDSC registration (oryon_pointdsc_pose) against the golden fixtures written by the
unmodified reference (models/pointdsc/PointDSC.py, run by oracle/make_golden.py) and against the CPU
oracle on further seeded cases.

Tolerances (float32 network evaluated with a different summation order than ATen's):
  confidence logits            2e-4 abs/rel
  seeds                        identical rank order of seed scores; identical indices wherever the score is
                               unique.  score = conf * is_local_max is exactly 0 for every suppressed point, and
                               ATen's argsort(descending=True) is not stable, so WHICH zero-score points fill the
                               tail of the seed list is implementation-defined in the reference itself (CPU AVX
                               sort vs CUDA sort differ); the library orders ties by ascending index.
  final transform              1e-4 abs on every entry of the 4x4 (SURVEY.md 8c gate)
"""
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from gpu_util import need_gpu
from oryon_b200 import synth
from oryon_b200.utils.pointdsc import init as pdsc

pytestmark = pytest.mark.gpu

CFG = synth.POINTDSC_DEFAULT_CFG


POSE_TOL = {"fp32": 1e-4, "tcgen05": 2e-4}


@pytest.fixture(autouse=True, params=["fp32", "tcgen05"])
def network_path(request, monkeypatch):
    """Every test runs on both forms of the NonLocal network: the default fp32 CUDA-core layer kernel, held to the 1e-4 pose gate of
    SURVEY.md 8(c), and the opt-in form on the tcgen05 GEMM (ORYON_PDSC_TC=1, read per call; three fp16 products per product with the
    tensor core's truncating fp32 accumulation), which reaches 1.04e-4 on the hardest case and is held to 2e-4 -- the reason it is
    not the default."""
    if request.param == "tcgen05":
        monkeypatch.setenv("ORYON_PDSC_TC", "1")
    else:
        monkeypatch.delenv("ORYON_PDSC_TC", raising=False)
    return request.param


def _check_seeds(conf_ref, src, mine, theirs, radius):
    """Seed scores (PointDSC.py:211-217) from the reference confidences; compare rank-ordered scores and indices."""
    conf_ref = torch.as_tensor(conf_ref)
    d = torch.norm(src[:, None, :] - src[None, :, :], dim=-1)
    rel = (conf_ref[:, None] >= conf_ref[None, :]) | (d >= radius)
    val = (conf_ref * rel.min(-1)[0].float()).numpy()
    mine, theirs = np.asarray(mine, dtype=np.int64), np.asarray(theirs, dtype=np.int64)
    assert mine.shape == theirs.shape and len(set(mine.tolist())) == len(mine) and mine.min() >= 0
    assert np.array_equal(val[mine], val[theirs]), "seed scores differ in rank order"
    uniq = np.array([np.sum(val == v) == 1 for v in val[theirs]])
    assert np.array_equal(mine[uniq], theirs[uniq]), "seed indices differ where the score is unique"
    tied = ~uniq
    if tied.any():  # library rule: ascending index among equal scores
        for v in np.unique(val[mine[tied]]):
            grp = mine[val[mine] == v]
            assert np.all(np.diff(grp) > 0)


def _solver(seed):
    sd = synth.pointdsc_state_dict(seed)
    return sd, pdsc.PointDSCSolver(sd, in_dim=CFG["in_dim"], num_layers=CFG["num_layers"], num_channels=CFG["num_channels"],
                                   num_iterations=CFG["num_iterations"], ratio=CFG["ratio"], sigma_d=CFG["sigma_d"], k=CFG["k"],
                                   nms_radius=CFG["inlier_threshold"], device="cuda:0")


@pytest.mark.parametrize("seed", list(synth.POINTDSC_CASES))
def test_pointdsc_matches_reference_golden(golden_dir, seed, network_path):
    need_gpu()
    g = np.load(os.path.join(golden_dir, f"pointdsc_{seed}.npz"))
    n, out_frac = synth.POINTDSC_CASES[seed]
    sd, solver = _solver(seed)
    data = synth.rigid_correspondences(seed, n=n, outlier_frac=out_frac)
    assert [synth.tensor_checksum(t) for t in (data["src"], data["tgt"], sd["encoder.layer0.weight"])] == list(g["in_sum"])
    T, dbg = pdsc.pointdsc_poses(solver, [data["src"]], [data["tgt"]], return_debug=True)
    np.testing.assert_allclose(dbg["conf"][0, :n].cpu().numpy(), g["conf"], rtol=2e-4, atol=2e-4)
    _check_seeds(g["conf"], data["src"], dbg["seeds"][0, :len(g["seeds"])].cpu().numpy(), g["seeds"], CFG["inlier_threshold"])
    np.testing.assert_allclose(T[0].cpu().numpy(), g["final_trans"], atol=POSE_TOL[network_path])
    # reference interface: [4,4] float32 CPU tensor
    single = pdsc.get_pointdsc_pose(solver, data["src"], data["tgt"], "cuda:0")
    assert single.device.type == "cpu" and single.dtype == torch.float32 and single.shape == (4, 4)
    assert torch.equal(single, T[0].cpu())
    # a rigid motion: orthonormal rotation with det +1, last row 0 0 0 1
    R = single[:3, :3].double()
    assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-5) and abs(torch.det(R).item() - 1) < 1e-5
    assert torch.equal(single[3], torch.tensor([0.0, 0.0, 0.0, 1.0]))


def test_pointdsc_batched_equals_single_and_oracle(network_path):
    """Ragged batch (different n per pair) in one call == per-pair calls == CPU oracle."""
    need_gpu()
    sd, solver = _solver(300)
    cases = [(400, 500, 0.3), (401, 333, 0.5), (402, 64, 0.1), (403, 500, 0.45), (404, 211, 0.0)]
    datas = [synth.rigid_correspondences(s, n=n, outlier_frac=f) for s, n, f in cases]
    T, dbg = pdsc.pointdsc_poses(solver, [d["src"] for d in datas], [d["tgt"] for d in datas], return_debug=True)
    torch.set_num_threads(8)
    for p, d in enumerate(datas):
        one = pdsc.pointdsc_poses(solver, [d["src"]], [d["tgt"]])
        assert torch.equal(one[0], T[p]), f"pair {p}: batched result differs from the single call"
        ref, rdbg = oracle.pointdsc_pose(sd, CFG, d["src"], d["tgt"], return_debug=True)
        n = d["src"].shape[0]
        np.testing.assert_allclose(dbg["conf"][p, :n].cpu().numpy(), rdbg["conf"].numpy(), rtol=2e-4, atol=2e-4)
        ns = rdbg["seeds"].shape[0]
        _check_seeds(rdbg["conf"], d["src"], dbg["seeds"][p, :ns].cpu().numpy(), rdbg["seeds"].numpy(), CFG["inlier_threshold"])
        np.testing.assert_allclose(T[p].cpu().numpy(), ref.numpy(), atol=POSE_TOL[network_path])


def test_pointdsc_duplicate_correspondences(network_path):
    """Sampling with replacement (utils/misc.py:242-254, n > N) feeds duplicated rows: exact score and
    distance ties must not change the pose."""
    need_gpu()
    sd, solver = _solver(301)
    d = synth.rigid_correspondences(410, n=120, outlier_frac=0.2)
    g = torch.Generator().manual_seed(5)
    sel = torch.randint(0, 120, (500,), generator=g)
    src, tgt = d["src"][sel], d["tgt"][sel]
    T = pdsc.pointdsc_poses(solver, [src], [tgt])[0].cpu()
    torch.set_num_threads(8)
    ref = oracle.pointdsc_pose(sd, CFG, src, tgt)
    np.testing.assert_allclose(T.numpy(), ref.numpy(), atol=POSE_TOL[network_path])
    np.testing.assert_allclose(T.numpy(), d["T"].numpy(), atol=5e-3)  # and it is the planted motion


def test_pointdsc_recovers_planted_motion_full_size():
    """Size-independent property at the reference's working size (500 correspondences, 32 pairs at once):
    the planted rigid motion is recovered."""
    need_gpu()
    _, solver = _solver(302)
    datas = [synth.rigid_correspondences(500 + i, n=500, outlier_frac=0.3) for i in range(32)]
    T = pdsc.pointdsc_poses(solver, [d["src"] for d in datas], [d["tgt"] for d in datas]).cpu()
    for i, d in enumerate(datas):
        assert (T[i] - d["T"]).abs().max().item() < 5e-3, i


def test_pointdsc_errors():
    need_gpu()
    _, solver = _solver(300)
    from oryon_b200 import _lib
    few = torch.rand(5, 3)
    with pytest.raises(_lib.OryonError):  # int(5 * 0.1) == 0 seeds: the reference raises too
        pdsc.pointdsc_poses(solver, [few], [few])
    with pytest.raises(ValueError):
        pdsc.pointdsc_poses(solver, [torch.rand(20, 3)], [torch.rand(21, 3)])
