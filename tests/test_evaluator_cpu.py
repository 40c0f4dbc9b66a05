"""N2/N3 on the CPU: (1) the evaluator oracle (oracle/eval_oracle.py) against raw error values produced by the reference's
own functions (tests/golden/eval_0.npz, oracle/make_golden_eval.py); (2) the bookkeeping of the Evaluator mirror
(oryon_b200/utils/evaluator.py) -- metric lists, counts, means, LaTeX row -- against the reference Evaluator's state
(tests/golden/eval_0.json), with the oracle injected as the error backend (the product default is the CUDA library,
covered by tests/test_evaluator_gpu.py); (3) the prediction-CSV round trip."""
import json
import os

import numpy as np
import pytest
import torch

import eval_oracle
from oryon_b200 import synth
from oryon_b200.utils.evaluator import Evaluator, dict_from_preds, format_sym_set

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def data():
    obj, cs = synth.eval_objects(0), synth.eval_cases(0)
    g = np.load(os.path.join(GOLDEN, "eval_0.npz"))
    assert list(g["in_sum"]) == [synth.tensor_checksum(cs["pred_pose"]), synth.tensor_checksum(cs["gt_pose"]),
                                 synth.tensor_checksum(torch.from_numpy(obj["models"][2]["pts"]))], "synthetic inputs drifted"
    return obj, cs, g["raw"], json.load(open(os.path.join(GOLDEN, "eval_0.json")))


def _effective_pred(cs, i):
    pred = cs["pred_pose"][i].numpy().copy()
    return np.eye(4, dtype=pred.dtype) if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1 else pred


def test_oracle_matches_reference_values(data):
    obj, cs, raw, _ = data
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    for i, cid in enumerate(cs["cls_id"]):
        if np.isnan(raw[i, 0]):
            continue
        pred, gt = _effective_pred(cs, i), cs["gt_pose"][i].numpy()
        r, t = eval_oracle.rt_errors(pred, gt)
        np.testing.assert_allclose([r, t], raw[i, :2], rtol=1e-9, atol=1e-9)
        pts = obj["models"][cid]["pts"]
        if syms[cid].shape[0] > 1:
            np.testing.assert_allclose(eval_oracle.adds_error(pts / 1000., pred, gt), raw[i, 2], rtol=1e-12)
        else:
            assert eval_oracle.add_error(pts / 1000., pred, gt) == raw[i, 2]          # float16 value, bit-exact
        ms, mp = eval_oracle.mssd_mspd(pts, syms[cid], pred, gt, cs["camera"].numpy())
        np.testing.assert_allclose([ms, mp], raw[i, 4:6], rtol=1e-9)


def _run_mirror(obj, cs, backend, batched):
    ev = Evaluator("synthetic", compute_vsd=False, compute_iou=True, pose_errors=backend)
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    n = len(cs["cls_id"])

    def res(idx):
        sl = torch.tensor(idx)
        return {"iou_a": cs["iou_a"][sl], "iou_q": cs["iou_q"][sl], "gt_pose": cs["gt_pose"][sl], "pred_pose": cs["pred_pose"][sl],
                "pred_pose_rel": cs["pred_pose_rel"][sl], "cls_id": [cs["cls_id"][i] for i in idx],
                "camera": [cs["camera"].numpy() for _ in idx], "depth": [None for _ in idx],
                "instance_id": [cs["instance_id"][i] for i in idx]}

    fail = {"iou_a": cs["iou_a"][7:8], "iou_q": cs["iou_q"][7:8], "cls_id": [cs["cls_id"][7]], "instance_id": [cs["instance_id"][7]]}
    if batched:      # whole batches, the failure in its place
        ev.register_test(res(list(range(0, 7))))
        ev.register_test_failure(fail)
        ev.register_test(res(list(range(8, n))))
    else:
        for i in range(n):
            ev.register_test_failure(fail) if i == 7 else ev.register_test(res([i]))
    return ev


@pytest.mark.parametrize("batched", [False, True])
def test_evaluator_bookkeeping_matches_reference(data, batched):
    obj, cs, _, gold = data
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    backend = lambda cls_ids, pred, gt, cams: eval_oracle.pose_errors(obj["models"], syms, cls_ids, pred, gt, cams)  # noqa: E731
    ev = _run_mirror(obj, cs, backend, batched)
    for k, v in gold["metrics"].items():
        if k == "instance_id":
            assert ev.metrics[k] == v
        elif k in ("R error", "T error"):
            np.testing.assert_allclose(ev.metrics[k], v, rtol=1e-9, atol=1e-9)
        else:
            assert [float(x) for x in ev.metrics[k]] == v, k
    assert {k: [int(x) for x in v] for k, v in ev.counts.items()} == gold["counts"]
    means = ev.get_means()
    for k, v in gold["means"].items():
        np.testing.assert_allclose(means[k], v, rtol=1e-9)
    for c in (1, 2, 3):
        om = ev.get_obj_means(c)
        for k, v in gold["obj_means"][str(c)].items():
            np.testing.assert_allclose(om[k], v, rtol=1e-9)
    assert ev.get_latex_str() == gold["latex"]


class _OracleBackend:
    """Error backend for CPU tests: the oracles in place of the CUDA library."""

    def __init__(self, models, syms):
        self.models, self.syms = models, syms

    def __call__(self, cls_ids, pred, gt, cams):
        return eval_oracle.pose_errors(self.models, self.syms, cls_ids, pred, gt, cams)

    def vsd(self, cls_ids, pred, gt, cams, depths, diameters, delta, taus):
        import vsd_oracle
        out = np.zeros((len(cls_ids), len(taus)))
        for i, cid in enumerate(cls_ids):
            m, K = self.models[cid], np.asarray(cams[i]).reshape(3, 3)
            H, W = np.asarray(depths[i]).shape

            def render(T):
                p16 = np.asarray(T).astype(np.float16)
                R, t = p16[:3, :3].astype(np.float32), (p16[:3, 3] * np.float16(1000)).astype(np.float16).astype(np.float32)
                return vsd_oracle.rasterize_depth(m["pts"], m["faces"], R, t, K[0, 0], K[1, 1], K[0, 2], K[1, 2], H, W)

            out[i] = vsd_oracle.vsd_errors(render(pred[i]), render(gt[i]), np.asarray(depths[i]), K, delta, taus, diameters[i])
        return out


def _vsd_inputs():
    import vsd_oracle
    obj, cs = synth.eval_mesh_objects(0), synth.eval_cases(1, n=9)
    K = cs["camera"].numpy()
    depths = [synth.eval_scene_depth(obj["models"], cs["cls_id"][i], cs["gt_pose"][i].numpy(), K, i, render=vsd_oracle.rasterize_depth)
              for i in range(9)]
    g = np.load(os.path.join(GOLDEN, "vsd_0.npz"))
    assert list(g["depth_sum"]) == [synth.tensor_checksum(torch.from_numpy(d)) for d in depths], "synthetic scenes drifted"
    return obj, cs, K, depths, g["errs"], json.load(open(os.path.join(GOLDEN, "vsd_0.json")))


def test_vsd_oracle_and_evaluator_match_reference():
    """VSD errors of the oracle == the reference's vsd() on the same rendered depth images (tests/golden/vsd_0.npz), and the
    Evaluator mirror with compute_vsd=True reproduces the reference Evaluator's VSD / AR / LaTeX row (vsd_0.json)."""
    obj, cs, K, depths, errs, gold = _vsd_inputs()
    syms = {k: format_sym_set(s) for k, s in obj["syms"].items()}
    be = _OracleBackend(obj["models"], syms)
    ev = Evaluator("synthetic", compute_vsd=True, compute_iou=True, pose_errors=be)
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    idx = list(range(9))
    sl = torch.tensor(idx)
    ev.register_test({"iou_a": cs["iou_a"][sl], "iou_q": cs["iou_q"][sl], "gt_pose": cs["gt_pose"][sl], "pred_pose": cs["pred_pose"][sl],
                      "pred_pose_rel": cs["pred_pose_rel"][sl], "cls_id": list(cs["cls_id"]), "camera": [K] * 9, "depth": depths,
                      "instance_id": list(cs["instance_id"])})
    pred = cs["pred_pose"].numpy().astype(np.float64)
    for i in range(9):
        if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1:
            pred[i] = np.eye(4)
    got = be.vsd(cs["cls_id"], pred, cs["gt_pose"].numpy(), np.stack([K] * 9), depths, [obj["diams"][c] for c in cs["cls_id"]], 15.,
                 list(np.arange(0.05, 0.51, 0.05)))
    np.testing.assert_array_equal(got, errs)                       # integer pixel counts: exact
    for k in ("VSD", "AR", "MSSD", "MSPD", "ADD(S)-0.1d"):
        assert [float(x) for x in ev.metrics[k]] == gold["metrics"][k], k
    assert list(ev.metrics.keys()) == list(gold["metrics"].keys())   # same key order -> same JSON from save()
    assert ev.get_latex_str() == gold["latex"]
