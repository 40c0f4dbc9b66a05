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


def test_vsd_is_loud():
    with pytest.raises(NotImplementedError):
        Evaluator("x", compute_vsd=True)


def test_prediction_csv_round_trip(tmp_path):
    from oryon_b200.pipeline import FPM_Pipeline
    pipe = FPM_Pipeline.__new__(FPM_Pipeline)
    path = tmp_path / "pred.csv"
    pipe.pred_file = open(path, "w")
    rng = np.random.RandomState(0)
    poses = [np.vstack([rng.randn(3, 4).astype(np.float32), [[0, 0, 0, 1]]]) for _ in range(3)]
    for i, p in enumerate(poses):
        pipe.add_pred_pose(f"scene{i} {10 + i} mug", f"scene{i} {20 + i} mug", np.float32(0.5 + 0.1 * i), np.float32(0.25), p)
    pipe.pred_file.close()
    preds, ia, iq, present = dict_from_preds(str(path))
    assert present and len(preds) == 3
    for i, p in enumerate(poses):
        key = f"scene{i}_{10 + i}_scene{i}_{20 + i}_mug"
        np.testing.assert_array_equal(preds[key].astype(np.float32), p[:3].astype(np.float32))   # str(float32) round-trips
        assert np.float32(ia[key]) == np.float32(0.5 + 0.1 * i) and iq[key] == 0.25      # shortest repr of a float32 round-trips
    three = tmp_path / "three.csv"
    three.write_text("s 1 mug,s 2 mug," + " ".join(["1.0"] * 12) + "\n")
    assert dict_from_preds(str(three))[3] is False
    bad = tmp_path / "bad.csv"
    bad.write_text("a,b\n")
    with pytest.raises(RuntimeError):
        dict_from_preds(str(bad))
