#!/usr/bin/env python
"""Offline scorer: the B200 counterpart of the reference's ``scripts/evaluation/compute_metrics.py`` (:52-129).

    python scripts/evaluation/compute_metrics.py predictions.csv --dataset-type nocs --root data --dataset nocs \\
           --split cross_scene_test --obj all [--mask predicted] [--no-vsd] [--out table.tex]

Reads a prediction CSV (``id_a,id_q,<12 floats>[,iou_a,iou_q]``, written by ``FPM_Pipeline.add_pred_pose`` /
``run_test.py``), walks the dataset's pair split, and for every pair registers with the evaluator exactly what the
reference's scorer registers (:88-115): the predicted query pose ``pred_rel @ gt_anchor``, the ground-truth query pose, the
relative prediction, object id, pair id, the query intrinsics and depth frame (+ the IoUs when the CSV has them); pairs the
dataset marks invalid are registered as failures.  The pose errors (ADD(S), MSSD, MSPD, VSD, AR) are evaluated on the GPU by
``oryon_b200.utils.evaluator.Evaluator``; the metrics JSON is written next to the CSV and the LaTeX row printed / appended.

Differences from the reference script, all on the host side: the dataset is configured from the command line instead of the
``config_*.yaml`` hydra leaves next to the CSV (omegaconf is not a dependency); failure rows carry the IoUs when the CSV has
them (the reference's failure payload omits them although its evaluator reads them when ``compute_iou`` is set); the 1-based
OBJ face indices of the NOCS models are shifted to 0-based before they reach the rasteriser (the reference hands them to
OpenGL as they are).
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, Iterator, Optional, Tuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oryon_b200.utils.evaluator import Evaluator, dict_from_preds, zero_based_faces  # noqa: E402


def pair_results(dataset, preds: dict, ious_a: dict, ious_q: dict, iou_present: bool) -> Iterator[Tuple[bool, dict]]:
    """``(valid, payload)`` per pair of the split, in dataset order: the argument of ``Evaluator.register_test`` (valid) or
    ``register_test_failure`` (compute_metrics.py:88-115)."""
    for idx in range(len(dataset)):
        item_a, item_q, _, _, cls_id, instance_id, valid = dataset[idx]
        gt_q = np.asarray(item_q["metadata"]["poses"][0])
        gt_a = np.asarray(item_a["metadata"]["poses"][0])
        pred_rel = np.concatenate([preds[instance_id], np.asarray([[0., 0., 0., 1.]])], axis=0)
        pred_q = pred_rel @ gt_a
        if valid:
            res = {"gt_pose": torch.tensor(gt_q).unsqueeze(0), "pred_pose": torch.tensor(pred_q).unsqueeze(0),
                   "pred_pose_rel": torch.tensor(pred_rel).unsqueeze(0), "cls_id": [cls_id], "instance_id": [instance_id],
                   "camera": [np.asarray(item_q["camera"])], "depth": [np.asarray(item_q["depth"]).squeeze()]}
        else:
            res = {"cls_id": [cls_id], "instance_id": [instance_id]}
        if iou_present:
            res["iou_a"] = torch.tensor(ious_a[instance_id]).unsqueeze(0)
            res["iou_q"] = torch.tensor(ious_q[instance_id]).unsqueeze(0)
        yield bool(valid), res


def compute_metrics(results_file: str, dataset, exp_tag: str = "", compute_vsd: bool = True, print_summary: bool = False,
                    out_file: Optional[str] = None, pose_errors: Optional[Callable] = None, failed: Optional[set] = None) -> Evaluator:
    """``failed``: pair ids (``<scene_a>_<img_a>_<scene_q>_<img_q>_<obj>``) whose IN-LOOP status was not 'ok' (no correspondences,
    empty predicted mask).  The reference's offline scorer cannot know them -- it judges validity from the dataset alone and scores
    the identity pose such pairs carry in the CSV -- but its test loop registers them with ``register_test_failure``
    (pipeline.py:335-342), which zeroes every metric of the pair.  Passing the set reproduces the evaluator state the reference's
    ``on_test_end`` saves; leaving it ``None`` reproduces scripts/evaluation/compute_metrics.py."""
    metric_file = os.path.splitext(results_file)[0] + ".json"
    preds, ious_a, ious_q, iou_present = dict_from_preds(results_file)
    if not iou_present:
        print(f"IoU not found in prediction file {results_file}.")
    evaluator = Evaluator(exp_tag=exp_tag, compute_vsd=compute_vsd, compute_iou=iou_present, pose_errors=pose_errors)
    evaluator.init_test()
    models, diams, symms = dataset.get_object_info()
    evaluator.add_object_info(zero_based_faces(models), diams, symms)
    for valid, res in pair_results(dataset, preds, ious_a, ious_q, iou_present):
        if valid and not (failed and res["instance_id"][0] in failed):
            evaluator.register_test(res)
        else:
            evaluator.register_test_failure(res)
    if print_summary:
        print(evaluator.test_summary())
    latex = evaluator.get_latex_str()
    if out_file is None:
        print(latex)
    else:
        with open(out_file, "a") as f:
            f.write(latex)
    with open(metric_file, "w") as f:
        evaluator.save(f)
    return evaluator


def main(argv=None):
    ap = argparse.ArgumentParser(description="Score a prediction CSV against a dataset pair split")
    ap.add_argument("results", help="prediction CSV")
    ap.add_argument("--dataset-type", default=None, choices=["nocs", "toyl"], help="default: 'nocs' / 'toyl' found in the CSV path, as the reference does")
    ap.add_argument("--root", default="data")
    ap.add_argument("--dataset", default=None, help="dataset folder under --root (default: the dataset type)")
    ap.add_argument("--split", default="cross_scene_test")
    ap.add_argument("--obj", default="all")
    ap.add_argument("--mask", default="predicted")
    ap.add_argument("--exp-tag", default="")
    ap.add_argument("--no-vsd", action="store_true", help="skip VSD / AR (no depth rasterisation)")
    ap.add_argument("--summary", action="store_true")
    ap.add_argument("--out", default=None, help="append the LaTeX row to this file instead of printing it")
    args = ap.parse_args(argv)
    from oryon_b200.datasets import NOCSDataset, TOYLDataset
    kind = args.dataset_type or ("nocs" if "nocs" in args.results else "toyl" if "toyl" in args.results else None)
    if kind is None:
        raise RuntimeError("Dataset not supported")
    cfg = dict(dataset=dict(root=args.root, max_corrs=500, img_size=[224, 224], test=dict(name=args.dataset or kind, split=args.split, obj=args.obj)),
               test=dict(mask=args.mask, add_description="yes"))
    dataset = {"nocs": NOCSDataset, "toyl": TOYLDataset}[kind](cfg, eval=True)
    compute_metrics(args.results, dataset, args.exp_tag, not args.no_vsd, args.summary, args.out)


if __name__ == "__main__":
    main()
