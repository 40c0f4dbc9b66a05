"""Generate ``tests/golden/scorer_0.json``: what the reference's OWN offline scorer (scripts/evaluation/compute_metrics.py:52-129,
``compute_metrics`` and ``dict_from_preds`` exec'd from the source text as it lies under /root/reference) writes for a
prediction CSV over the synthetic TOYL tree of ``oryon_b200.synth.write_toyl_tree`` at 480 x 640 -- the metrics JSON of
``Evaluator.save`` and the LaTeX row -- with the reference's ``TOYLDataset`` and ``Evaluator(compute_vsd=True)`` underneath.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_scorer.py

Environment notes.  Everything ``make_golden_toyl.py`` needs (matplotlib / pytz / plyfile stand-ins, ``F.resize`` without
antialiasing).  ``bop_toolkit_lib.renderer_vispy`` is replaced by the oracle's software rasteriser as in
``make_golden_vsd.py`` (OpenGL cannot run here): the frames are written at the 480 x 640 the reference hard-codes for its
renderer (evaluator.py:97).  ``OmegaConf.load`` returns the configuration built here instead of reading the
``config_*.yaml`` hydra leaf next to the CSV; ``open_dict`` is a no-op context.  TOYL rather than NOCS because the NOCS
models carry 1-based OBJ face indices, which the reference hands to OpenGL unshifted (an out-of-range vertex in the stand-in).

The predictions: the ground-truth relative pose ``gt_q @ inv(gt_a)`` of every pair, perturbed by a seeded rotation /
translation of growing size (the last pair gets an all-zero pose = the pipeline's failure output), written in the wire format
of ``FPM_Pipeline.add_pred_pose`` with float32 values and IoUs.  The CSV text is part of the fixture.
"""
import contextlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import make_golden_toyl as mgt  # noqa: E402  (installs the shims and stand-ins at import)
import make_golden_nocs as mgn  # noqa: E402
import vsd_oracle  # noqa: E402

_stub = types.ModuleType("bop_toolkit_lib.renderer_vispy")
_stub.RendererVispy = vsd_oracle.OracleRenderer
sys.modules["bop_toolkit_lib.renderer_vispy"] = _stub

import torch  # noqa: E402
from typing import Optional  # noqa: E402
from utils.evaluator import Evaluator  # noqa: E402  (reference)

from oryon_b200 import synth  # noqa: E402

HW = (480, 640)


def reference_scorer(datasets_env: dict, args, evaluator_cls=Evaluator) -> dict:
    src = open(os.path.join(mgt.ref_shims.REFERENCE_ROOT, "scripts", "evaluation", "compute_metrics.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def dict_from_preds"))
    end = next(i for i, l in enumerate(src) if l.startswith("def main"))

    class OmegaConf:
        load = staticmethod(lambda path: args)

    env = {"os": os, "sys": sys, "np": np, "json": json, "torch": torch, "Optional": Optional, "join": os.path.join, "OmegaConf": OmegaConf,
           "open_dict": lambda a: contextlib.nullcontext(a), "Evaluator": evaluator_cls, "NOCSDataset": datasets_env.get("NOCSDataset"),
           "TOYLDataset": datasets_env.get("TOYLDataset")}
    exec("\n".join(src[start:end]), env)
    return env


def perturbed_predictions(ds, seed: int):
    """CSV lines for every pair of the reference dataset ``ds`` (see the module docstring)."""
    g = np.random.default_rng(77 + seed)
    lines = []
    for i in range(len(ds)):
        item_a, item_q, _, _, _, _, cls_id, instance_id, _ = ds[i]
        gt_a, gt_q = item_a["metadata"]["poses"][0].numpy().astype(np.float64), item_q["metadata"]["poses"][0].numpy().astype(np.float64)
        rel = gt_q @ np.linalg.inv(gt_a)
        d = np.eye(4)
        d[:3, :3] = synth._axis_rotation(g.normal(size=3), 0.02 * (i + 1) ** 2)
        d[:3, 3] = g.normal(size=3) * 0.002 * (i + 1) ** 2
        pred = (d @ rel).astype(np.float32) if i != len(ds) - 1 else np.zeros((4, 4), np.float32)
        sa, ia, sq, iq, obj = instance_id.split("_", 4)          # NOCS object names hold underscores
        pose = " ".join(str(n) for n in pred[:3, :].flatten())
        lines.append(",".join([f"{sa} {ia} {obj}", f"{sq} {iq} {obj}", pose, str(np.float32(g.uniform(0.2, 1.0))), str(np.float32(g.uniform(0.2, 1.0)))]) + "\n")
    return lines


def main(seed=0):
    env = mgt.reference_dataset_classes()
    with tempfile.TemporaryDirectory() as d:
        info = synth.write_toyl_tree(d, seed, hw=HW)
        args = mgt.cfg(dict(augs=dict(), debug_valid="no", use_seed=False, seed=1, exp_tag="synthetic",
                            dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj="all")),
                            test=dict(mask="predicted", add_description="yes")))
        lines = perturbed_predictions(env["TOYLDataset"](args, eval=True), seed)
        out = {}
        # 'all' holds a pair without ground-truth correspondences -> a failure row; the reference's failure payload carries no IoUs
        # (:111-114) although its evaluator reads them when the CSV has them (evaluator.py:311): with IoUs only an all-valid split runs
        for tag, obj_split, csv_lines in (("noiou", "all", [",".join(l.split(",")[:3]) + "\n" for l in lines]), ("iou", "ducks", lines)):
            args["dataset"]["test"]["obj"] = obj_split
            os.makedirs(os.path.join(d, "results"), exist_ok=True)
            csv = os.path.join(d, "results", f"toyl_{tag}_r_s_t.csv")      # 'toyl' in the path selects the dataset class (:78-81)
            open(csv, "w").writelines(csv_lines)
            scorer = reference_scorer(env, args)
            tex = os.path.join(d, "results", f"{tag}.tex")
            scorer["compute_metrics"](csv, False, tex)
            out[tag] = dict(obj=obj_split, csv=csv_lines, metrics=json.load(open(os.path.splitext(csv)[0] + ".json")), latex=open(tex).read())
            print(tag, out[tag]["latex"])
    with open(os.path.join(ROOT, "tests", "golden", f"scorer_{seed}.json"), "w") as f:
        json.dump(out, f, indent=1)
    main_nocs(seed)


class EvaluatorWithoutVSD(Evaluator):
    """The reference evaluator with ``compute_vsd`` forced off: the scorer hard-codes ``compute_vsd=True`` (:84), and the NOCS
    models carry 1-based OBJ face indices that only the reference's OpenGL renderer can take as they are."""

    def __init__(self, exp_tag, compute_vsd=True, compute_iou=True):
        super().__init__(exp_tag, compute_vsd=False, compute_iou=compute_iou)


def main_nocs(seed=0):
    """``tests/golden/scorer_nocs_<seed>.json``: the same for the NOCS tree (pairs keyed by object NAME), VSD / AR off."""
    env = mgn.reference_dataset_classes()
    with tempfile.TemporaryDirectory() as d:
        info = synth.write_nocs_tree(d, seed)
        args = mgt.cfg(dict(augs=dict(), debug_valid="no", use_seed=False, seed=1, exp_tag="synthetic nocs",
                            dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj="all")),
                            test=dict(mask="predicted", add_description="yes")))
        lines = perturbed_predictions(env["NOCSDataset"](args, eval=True), seed + 1)
        csv_lines = [",".join(l.split(",")[:3]) + "\n" for l in lines]          # the split holds an invalid pair: no IoUs (see main)
        os.makedirs(os.path.join(d, "results"), exist_ok=True)
        csv = os.path.join(d, "results", "nocs_noiou_r_s_t.csv")
        open(csv, "w").writelines(csv_lines)
        tex = os.path.join(d, "results", "nocs.tex")
        reference_scorer(env, args, EvaluatorWithoutVSD)["compute_metrics"](csv, False, tex)
        out = dict(obj="all", csv=csv_lines, metrics=json.load(open(os.path.splitext(csv)[0] + ".json")), latex=open(tex).read())
        print("nocs", out["latex"])
    with open(os.path.join(ROOT, "tests", "golden", f"scorer_nocs_{seed}.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
