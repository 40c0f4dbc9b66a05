"""Generate ``tests/golden/vsd_0.npz`` / ``vsd_0.json``: VSD errors and the VSD / AR entries of the evaluator, by running the
UNMODIFIED reference ``bop_toolkit_lib.pose_error.vsd`` and ``utils.evaluator.Evaluator(compute_vsd=True)``.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_vsd.py

The reference renders depth with OpenGL (bop_toolkit_lib/renderer_vispy.py), which cannot run here.  The module
``bop_toolkit_lib.renderer_vispy`` is therefore replaced by a stand-in whose ``RendererVispy`` is
``vsd_oracle.OracleRenderer`` (the oracle's software rasteriser): the reference code computes every VSD number from depth
images of that rasteriser.  What this pins is the error arithmetic (distance images, visibility masks, costs, recalls, AR);
the rendering step itself stays unpinned (see oracle/vsd_oracle.py).
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
import vsd_oracle  # noqa: E402

_stub = types.ModuleType("bop_toolkit_lib.renderer_vispy")
_stub.RendererVispy = vsd_oracle.OracleRenderer
sys.modules["bop_toolkit_lib.renderer_vispy"] = _stub

from utils.evaluator import Evaluator  # noqa: E402  (reference)
from bop_toolkit_lib.pose_error import vsd as ref_vsd  # noqa: E402

from oryon_b200 import synth  # noqa: E402


def main(seed=0, n=9):
    obj = synth.eval_mesh_objects(seed)
    cs = synth.eval_cases(seed + 1, n=n)
    K = cs["camera"].numpy()
    ev = Evaluator("synthetic", compute_vsd=True, compute_iou=True)
    ev.add_object_info(obj["models"], obj["diams"], obj["syms"])
    ev.init_test()
    taus = list(np.arange(0.05, 0.51, 0.05))
    errs, depths = np.zeros((n, len(taus))), []
    for i in range(n):
        cid = cs["cls_id"][i]
        depth = synth.eval_scene_depth(obj["models"], cid, cs["gt_pose"][i].numpy(), K, i, render=vsd_oracle.rasterize_depth)
        depths.append(depth)
        ev.register_test({"iou_a": cs["iou_a"][i:i + 1], "iou_q": cs["iou_q"][i:i + 1], "gt_pose": cs["gt_pose"][i:i + 1],
                          "pred_pose": cs["pred_pose"][i:i + 1], "pred_pose_rel": cs["pred_pose_rel"][i:i + 1], "cls_id": [cid],
                          "camera": [K], "depth": [depth], "instance_id": [cs["instance_id"][i]]})
        pred = cs["pred_pose"][i].numpy().copy()
        if np.count_nonzero(cs["pred_pose_rel"][i].numpy()) <= 1:
            pred = np.eye(4, dtype=pred.dtype)
        p16, g16 = pred.astype(np.float16), cs["gt_pose"][i].numpy().astype(np.float16)
        errs[i] = ref_vsd(p16[:3, :3], np.expand_dims(p16[:3, 3], axis=1) * 1000, g16[:3, :3], np.expand_dims(g16[:3, 3], axis=1) * 1000,
                          depth, K.reshape(3, 3), 15., taus, True, obj["diams"][cid], ev.renderer, cid)
    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, f"vsd_{seed}.npz"), errs=errs,
                        depth_sum=np.array([synth.tensor_checksum(torch.from_numpy(d)) for d in depths]),
                        in_sum=np.array([synth.tensor_checksum(cs["pred_pose"]), synth.tensor_checksum(torch.from_numpy(obj["models"][1]["pts"]))]))
    metrics = {k: [float(x) if not isinstance(x, str) else x for x in v] for k, v in ev.metrics.items()}
    with open(os.path.join(out, f"vsd_{seed}.json"), "w") as fh:
        json.dump(dict(metrics=metrics, means={k: float(v) for k, v in ev.get_means().items()}, latex=ev.get_latex_str()), fh)
    print(np.round(errs, 3))
    print({k: metrics[k] for k in ("VSD", "AR", "MSSD", "MSPD")})
    print(ev.get_latex_str())


if __name__ == "__main__":
    main()
