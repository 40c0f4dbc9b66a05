"""Generate ``tests/golden/eval_0.npz`` / ``eval_0.json`` by running the UNMODIFIED reference evaluator code
(utils/evaluator.py, utils/metrics.py, bop_toolkit_lib/pose_error.py) on the synthetic objects / poses of
``oryon_b200.synth.eval_objects`` / ``eval_cases``.

TEST INFRASTRUCTURE, build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_eval.py

``pytz`` (imported by bop_toolkit_lib/misc.py:9 for log timestamps only) is stubbed.  VSD / AR need the reference's
OpenGL renderer (bop_toolkit_lib/renderer_vispy.py), which cannot run here: the evaluator is driven with
``compute_vsd=False``.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

ref_shims.install()
sys.modules.setdefault("pytz", types.ModuleType("pytz"))

from utils.evaluator import Evaluator  # noqa: E402  (reference)
from utils.metrics import compute_add, compute_adds, compute_RT_distances  # noqa: E402
from utils.pcd import get_diameter  # noqa: E402
from bop_toolkit_lib.misc import format_sym_set  # noqa: E402
from bop_toolkit_lib.pose_error import my_mspd, my_mssd  # noqa: E402

from oryon_b200 import synth  # noqa: E402


def main(seed=0):
    obj = synth.eval_objects(seed)
    cs = synth.eval_cases(seed)
    n = len(cs["cls_id"])
    ev = Evaluator("synthetic", compute_vsd=False, compute_iou=True)
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    raw = np.zeros((n, 6))
    for i in range(n):
        cid = cs["cls_id"][i]
        if i == 7:   # a pair rejected before matching (pipeline.py:343-350)
            ev.register_test_failure({"iou_a": cs["iou_a"][i:i + 1], "iou_q": cs["iou_q"][i:i + 1], "cls_id": [cid],
                                      "instance_id": [cs["instance_id"][i]]})
            raw[i] = np.nan
            continue
        ev.register_test({"iou_a": cs["iou_a"][i:i + 1], "iou_q": cs["iou_q"][i:i + 1], "gt_pose": cs["gt_pose"][i:i + 1],
                          "pred_pose": cs["pred_pose"][i:i + 1], "pred_pose_rel": cs["pred_pose_rel"][i:i + 1], "cls_id": [cid],
                          "camera": [cs["camera"].numpy()], "depth": [None], "instance_id": [cs["instance_id"][i]]})
        # the raw error values behind the thresholded metrics, from the same reference functions (evaluator.py:229-277)
        pred = cs["pred_pose"][i].numpy().copy()
        if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1:
            pred = np.eye(4, dtype=pred.dtype)
        gt = cs["gt_pose"][i].numpy()
        model, sym = obj["models"][cid], format_sym_set(obj["syms"][cid])
        r, t = compute_RT_distances(pred, gt)
        add = compute_adds(model["pts"] / 1000., pred, gt) if sym.shape[0] > 1 else compute_add(model["pts"] / 1000., pred, gt)
        p16, g16 = pred.astype(np.float16), gt.astype(np.float16)
        pr, pt = p16[:3, :3], np.expand_dims(p16[:3, 3], axis=1) * 1000
        gr, gt_t = g16[:3, :3], np.expand_dims(g16[:3, 3], axis=1) * 1000
        raw[i] = [r[0], t[0], float(add), get_diameter(model["pts"]) / 1000.,
                  my_mssd(pr, pt, gr, gt_t, model["pts"], sym), my_mspd(pr, pt, gr, gt_t, cs["camera"].numpy(), model["pts"], sym)]
    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, f"eval_{seed}.npz"), raw=raw,
                        in_sum=np.array([synth.tensor_checksum(cs["pred_pose"]), synth.tensor_checksum(cs["gt_pose"]),
                                         synth.tensor_checksum(torch.from_numpy(obj["models"][2]["pts"]))]))
    metrics = {k: [float(x) if not isinstance(x, str) else x for x in v] for k, v in ev.metrics.items()}
    counts = {k: [int(x) for x in v] for k, v in ev.counts.items()}
    means = {k: float(v) for k, v in ev.get_means().items()}
    obj_means = {str(c): {k: float(v) for k, v in ev.get_obj_means(c).items()} for c in (1, 2, 3)}
    with open(os.path.join(out, f"eval_{seed}.json"), "w") as fh:
        json.dump(dict(metrics=metrics, counts=counts, means=means, obj_means=obj_means, latex=ev.get_latex_str()), fh)
    print(raw)
    print(means)
    print(ev.get_latex_str())


if __name__ == "__main__":
    main()
