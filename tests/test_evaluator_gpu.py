"""N2 on the GPU: ``oryon_eval_pose_errors`` (csrc/eval.cu) against the oracle and against the values the reference's
own functions produced (tests/golden/eval_0.npz), and the Evaluator mirror end to end on the CUDA backend against the
reference Evaluator's recorded state (tests/golden/eval_0.json)."""
import json
import os

import numpy as np
import pytest
import torch

import eval_oracle
from gpu_util import need_gpu
from oryon_b200 import synth
from oryon_b200.utils.evaluator import CudaPoseErrors, Evaluator, format_sym_set

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _inputs():
    obj, cs = synth.eval_objects(0), synth.eval_cases(0)
    n = len(cs["cls_id"])
    pred = cs["pred_pose"].numpy().astype(np.float64)
    for i in range(n):
        if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1:
            pred[i] = np.eye(4)
    cams = np.stack([cs["camera"].numpy()] * n)
    return obj, cs, pred, cs["gt_pose"].numpy(), cams


def _angle_tol_deg(theta_deg):
    # the reference normalises the float32 prediction in float32 (utils/metrics.py:251): ~1e-7 / sin(theta) rad of noise
    return np.degrees(4e-7 / np.maximum(np.sin(np.radians(theta_deg)), 1e-6)) + 1e-9


def test_pose_errors_match_oracle_and_reference():
    need_gpu()
    obj, cs, pred, gt, cams = _inputs()
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    be = CudaPoseErrors("cuda:0")
    for k, m in obj["models"].items():
        be.add_object(k, m["pts"], syms[k])
    got = be(cs["cls_id"], pred, gt, cams)
    ref = eval_oracle.pose_errors(obj["models"], syms, cs["cls_id"], pred, gt, cams)
    raw = np.load(os.path.join(GOLDEN, "eval_0.npz"))["raw"]
    for i, cid in enumerate(cs["cls_id"]):
        assert abs(got[i, 0] - ref[i, 0]) <= _angle_tol_deg(ref[i, 0]), (i, got[i, 0], ref[i, 0])
        np.testing.assert_allclose(got[i, 1], ref[i, 1], rtol=1e-12, atol=1e-12)
        assert got[i, 3] == float(syms[cid].shape[0] > 1)
        if syms[cid].shape[0] > 1:
            np.testing.assert_allclose(got[i, 2], ref[i, 2], rtol=1e-12)        # ADD-S: float64 mean of float64 distances
        else:
            assert abs(got[i, 2] - ref[i, 2]) <= np.spacing(np.float16(ref[i, 2])), (i, got[i, 2], ref[i, 2])   # float16 value
        np.testing.assert_allclose(got[i, 4:6], ref[i, 4:6], rtol=1e-11)
        if not np.isnan(raw[i, 0]):     # and straight against the reference's numbers
            assert abs(got[i, 0] - raw[i, 0]) <= _angle_tol_deg(raw[i, 0])
            np.testing.assert_allclose(got[i, 1], raw[i, 1], rtol=1e-9)
            np.testing.assert_allclose(got[i, 4:6], raw[i, 4:6], rtol=1e-9)
            if syms[cid].shape[0] > 1:
                np.testing.assert_allclose(got[i, 2], raw[i, 2], rtol=1e-12)
            else:
                assert abs(got[i, 2] - raw[i, 2]) <= np.spacing(np.float16(raw[i, 2]))
    # ADD on the asymmetric object is bit-identical on this fixture (the float16 emulation is exact up to the final mean)
    asym = [i for i, c in enumerate(cs["cls_id"]) if syms[c].shape[0] == 1]
    assert sum(got[i, 2] == ref[i, 2] for i in asym) >= len(asym) - 1


def test_large_model_and_many_poses():
    """A 6000-point symmetric model (several shared-memory sweeps, uneven chunks) and 70 poses (two launches)."""
    need_gpu()
    rng = np.random.RandomState(5)
    pts = (rng.rand(6001, 3) - 0.5) * 150.0
    sym = np.stack([np.hstack([synth._axis_rotation([0, 0, 1], 2 * np.pi * i / 5), np.zeros((3, 1))]) for i in range(5)])
    cs = synth.eval_cases(3, n=70)
    P = 70
    pred, gt = cs["pred_pose"].numpy().astype(np.float64), cs["gt_pose"].numpy()
    pred[5] = np.eye(4)
    cams = np.stack([cs["camera"].numpy()] * P)
    be = CudaPoseErrors("cuda:0")
    be.add_object("big", pts, sym)
    got = be(["big"] * P, pred, gt, cams)
    for i in (0, 5, 33, 69):
        np.testing.assert_allclose(got[i, 2], eval_oracle.adds_error(pts / 1000., pred[i], gt[i]), rtol=1e-12)
        ms, mp = eval_oracle.mssd_mspd(pts, sym, pred[i], gt[i], cams[i])
        np.testing.assert_allclose(got[i, 4:6], [ms, mp], rtol=1e-11)
    with pytest.raises(Exception):
        be2 = CudaPoseErrors("cuda:0")
        be2._ids["ghost"] = 12345
        be2(["ghost"], pred[:1], gt[:1], cams[:1])


def test_evaluator_mirror_on_cuda_matches_reference_state():
    need_gpu()
    obj, cs, _, _, _ = _inputs()
    gold = json.load(open(os.path.join(GOLDEN, "eval_0.json")))
    ev = Evaluator("synthetic", compute_vsd=False, compute_iou=True, device="cuda:0")
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    n = len(cs["cls_id"])

    def res(idx):
        sl = torch.tensor(idx)
        return {"iou_a": cs["iou_a"][sl], "iou_q": cs["iou_q"][sl], "gt_pose": cs["gt_pose"][sl], "pred_pose": cs["pred_pose"][sl],
                "pred_pose_rel": cs["pred_pose_rel"][sl], "cls_id": [cs["cls_id"][i] for i in idx],
                "camera": [cs["camera"].numpy() for _ in idx], "depth": [None for _ in idx],
                "instance_id": [cs["instance_id"][i] for i in idx]}

    ev.register_test(res(list(range(0, 7))))
    ev.register_test_failure({"iou_a": cs["iou_a"][7:8], "iou_q": cs["iou_q"][7:8], "cls_id": [cs["cls_id"][7]],
                              "instance_id": [cs["instance_id"][7]]})
    ev.register_test(res(list(range(8, n))))
    for k, v in gold["metrics"].items():
        if k == "instance_id":
            assert ev.metrics[k] == v
        elif k == "R error":
            assert all(abs(a - b) <= _angle_tol_deg(b) for a, b in zip(ev.metrics[k], v))
        elif k == "T error":
            np.testing.assert_allclose(ev.metrics[k], v, rtol=1e-9, atol=1e-9)
        else:
            assert [float(x) for x in ev.metrics[k]] == v, k          # every thresholded metric identical
    assert {k: [int(x) for x in v] for k, v in ev.counts.items()} == gold["counts"]
    assert ev.get_latex_str() == gold["latex"]


def _vsd_inputs():
    import vsd_oracle
    obj, cs = synth.eval_mesh_objects(0), synth.eval_cases(1, n=9)
    K = cs["camera"].numpy()
    depths = [synth.eval_scene_depth(obj["models"], cs["cls_id"][i], cs["gt_pose"][i].numpy(), K, i, render=vsd_oracle.rasterize_depth)
              for i in range(9)]
    g = np.load(os.path.join(GOLDEN, "vsd_0.npz"))
    assert list(g["depth_sum"]) == [synth.tensor_checksum(torch.from_numpy(d)) for d in depths], "synthetic scenes drifted"
    return obj, cs, K, depths, g["errs"], json.load(open(os.path.join(GOLDEN, "vsd_0.json")))


def test_vsd_matches_reference_arithmetic():
    """oryon_eval_vsd (CUDA rasteriser + visibility / cost reduction) == the reference's vsd() run on the oracle rasteriser's
    depth images (tests/golden/vsd_0.npz): the counts are integers, so the errors must be identical, which also shows the two
    rasterisers agree pixel for pixel on these scenes."""
    need_gpu()
    obj, cs, K, depths, errs, _ = _vsd_inputs()
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    be = CudaPoseErrors("cuda:0")
    for k, m in obj["models"].items():
        be.add_object(k, m["pts"], syms[k])
        be.add_mesh(k, m["faces"])
    pred = cs["pred_pose"].numpy().astype(np.float64)
    for i in range(9):
        if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1:
            pred[i] = np.eye(4)
    taus = list(np.arange(0.05, 0.51, 0.05))
    got = be.vsd(cs["cls_id"], pred, cs["gt_pose"].numpy(), np.stack([K] * 9), depths, [obj["diams"][c] for c in cs["cls_id"]], 15., taus)
    np.testing.assert_array_equal(got, errs)
    got_f = be.vsd(cs["cls_id"], pred, cs["gt_pose"].numpy(), np.stack([K] * 9), [d.astype(np.float32) for d in depths],
                   [obj["diams"][c] for c in cs["cls_id"]], 15., taus)
    np.testing.assert_array_equal(got_f, errs)                      # float32 test depth, same values
    with pytest.raises(Exception):                                   # an object without a mesh cannot be rendered
        be2 = CudaPoseErrors("cuda:0")
        be2.add_object("nomesh", obj["models"][1]["pts"], syms[1])
        be2.vsd(["nomesh"], pred[:1], cs["gt_pose"].numpy()[:1], K[None], depths[:1], [100.0], 15., taus)


def test_evaluator_with_vsd_on_cuda_matches_reference_state():
    need_gpu()
    obj, cs, K, depths, _, gold = _vsd_inputs()
    ev = Evaluator("synthetic", compute_vsd=True, compute_iou=True, device="cuda:0")
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    for lo, hi in ((0, 4), (4, 9)):                                  # two batches
        sl = torch.arange(lo, hi)
        ev.register_test({"iou_a": cs["iou_a"][sl], "iou_q": cs["iou_q"][sl], "gt_pose": cs["gt_pose"][sl], "pred_pose": cs["pred_pose"][sl],
                          "pred_pose_rel": cs["pred_pose_rel"][sl], "cls_id": cs["cls_id"][lo:hi], "camera": [K] * (hi - lo),
                          "depth": depths[lo:hi], "instance_id": cs["instance_id"][lo:hi]})
    for k in ("VSD", "AR", "MSSD", "MSPD", "ADD(S)-0.1d"):
        assert [float(x) for x in ev.metrics[k]] == gold["metrics"][k], k
    assert ev.get_latex_str() == gold["latex"]
    with pytest.raises(ValueError):
        bad = Evaluator("x", compute_vsd=True, device="cuda:0")
        bad.add_object_info(synth.eval_objects(0)["models"], synth.eval_objects(0)["diams"], synth.eval_objects(0)["syms"])
