"""CPU: the evaluator bookkeeping of ``FPM_Pipeline.test_step`` / ``validation_step`` (pipeline.py:321-350, :207-245) -- pair order,
batched successful runs, failure rows, the query ``eval_depth`` frame handed to the VSD term -- on a pipeline object built without
its GPU parts (``__new__``) and the evaluator on the oracle backend, against per-pair registration in the reference's order."""
import numpy as np
import torch

from oryon_b200 import synth
from oryon_b200.pipeline import FPM_Pipeline
from oryon_b200.utils.evaluator import Evaluator, format_sym_set
from test_evaluator_cpu import _OracleBackend


def _setup(compute_vsd):
    obj = synth.eval_mesh_objects(0)
    cs = synth.eval_cases(1, n=6)
    be = _OracleBackend(obj["models"], {k: format_sym_set(s) for k, s in obj["syms"].items()})

    def evaluator():
        ev = Evaluator("reg", compute_vsd=compute_vsd, compute_iou=True, pose_errors=be)
        ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
        ev.init_test()
        return ev
    return obj, cs, evaluator


def _batch_and_rows(cs, depths):
    n = len(cs["cls_id"])
    K = cs["camera"]
    batch = dict(cls_id=list(cs["cls_id"]), instance_id=list(cs["instance_id"]),
                 query=dict(pose=cs["gt_pose"], camera=torch.stack([K] * n), eval_depth=depths))
    status = ["ok", "ok", "invalid_mask", "ok", "no_corrs", "ok"]
    rows = [dict(status=status[i], iou_a=float(cs["iou_a"][i]), iou_q=float(cs["iou_q"][i]), pred_pose=cs["pred_pose"][i],
                 pred_pose_rel=cs["pred_pose_rel"][i]) for i in range(n)]
    return batch, rows, status


def test_register_matches_per_pair_registration_with_vsd():
    import vsd_oracle
    obj, cs, evaluator = _setup(True)
    K = cs["camera"].numpy()
    depths = torch.stack([torch.from_numpy(synth.eval_scene_depth(obj["models"], cs["cls_id"][i], cs["gt_pose"][i].numpy(), K, i, hw=(120, 160),
                                                                  render=vsd_oracle.rasterize_depth).astype(np.int32)) for i in range(6)])
    batch, rows, status = _batch_and_rows(cs, depths)
    pipe = FPM_Pipeline.__new__(FPM_Pipeline)
    pipe.evaluator = evaluator()
    pipe._register(batch, rows, test=True)
    ref = evaluator()                   # the reference's loop: one pair at a time, in order
    for i in range(6):
        if status[i] == "ok":
            ref.register_test({"iou_a": cs["iou_a"][i:i + 1], "iou_q": cs["iou_q"][i:i + 1], "gt_pose": cs["gt_pose"][i:i + 1],
                               "pred_pose": cs["pred_pose"][i:i + 1], "pred_pose_rel": cs["pred_pose_rel"][i:i + 1], "cls_id": [cs["cls_id"][i]],
                               "camera": [K], "depth": [depths[i].numpy()], "instance_id": [cs["instance_id"][i]]})
        else:
            ref.register_test_failure({"iou_a": cs["iou_a"][i:i + 1], "iou_q": cs["iou_q"][i:i + 1], "cls_id": [cs["cls_id"][i]],
                                       "instance_id": [cs["instance_id"][i]]})
    assert list(pipe.evaluator.metrics.keys()) == list(ref.metrics.keys())
    for k in ref.metrics:
        assert list(pipe.evaluator.metrics[k]) == list(ref.metrics[k]), k
    assert pipe.evaluator.counts == ref.counts and sum(ref.counts["Missing segm"]) == 2
    assert len(ref.metrics["VSD"]) == 6 and ref.metrics["instance_id"] == list(cs["instance_id"])


def test_register_validation_mode_keeps_no_instance_ids():
    obj, cs, evaluator = _setup(False)
    batch, rows, status = _batch_and_rows(cs, None)
    batch["query"].pop("eval_depth")
    pipe = FPM_Pipeline.__new__(FPM_Pipeline)
    pipe.evaluator = evaluator()
    pipe._register(batch, rows, test=False)
    ev = pipe.evaluator
    assert len(ev.metrics["R error"]) == 6 and ev.metrics["instance_id"] == [] and ev.metrics["cls_id"] == []
    assert ev.counts["Missing segm"] == [0, 0, 1, 0, 1, 0]
