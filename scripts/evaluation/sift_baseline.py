#!/usr/bin/env python
"""Key-point baseline: the B200 counterpart of the reference's ``scripts/evaluation/sift_nocs.py`` (:48-173) and
``sift_toyl.py`` (:54-190), two callers of the hot path that bypass the network.

    python scripts/evaluation/sift_baseline.py --dataset-type nocs|toyl --root data --pointdsc pretrained_models/pointdsc \\
           [--mask oracle|ovseg] [--mask-dir DIR] [--split cross_scene_test] [--obj all] [--no-vsd] [--out sift_nocs_oracle.txt]

Per pair of the dataset split: OpenCV SIFT key points and descriptors of the two grey frames (on the host, as in the
reference), key points outside the object mask dropped, then the path on the GPU -- descriptor matching
(``oryon_b200.utils.pcd.nn_correspondences_kp`` -> ``oryon_match_nn``; threshold 0.25, 500 rows), lifting of the matched key
points through the depth frames (``lift_pcd`` -> ``oryon_lift_pcd``, / 1000 to metres), PointDSC (``get_pointdsc_pose`` ->
``oryon_pointdsc_pose``) -- and the evaluator (``oryon_eval_pose_errors`` / ``oryon_eval_vsd``).  Pairs without a mask or
without key points inside it are registered as failures; every solved pair appends ``id_a,id_q,<12 floats>`` to the output
file (the three-token form of the prediction CSV).

Kept from the reference: the grey conversion uses OpenCV's BGR weights on the RGB frame (sift_nocs.py:85-86); key points are
cast to int16 (x, y); the NOCS variant matches all source descriptors and lets an empty match set raise, the TOYL variant
subsamples more than 1000 source descriptors and registers an empty match set as a failure; both lift with
``<reader>.get_camera()``, which for TOYL is the NOCS intrinsics (utils/data/toyl.py:19-21).  Differences, host side only: the
configuration comes from the command line (no hydra); ``--mask-dir`` replaces the hard-coded ``data/toyl/catseg_masks`` of the
``masks == 'ours'`` branch; the 1-based OBJ faces of the NOCS models are shifted before they reach the rasteriser.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oryon_b200.utils.evaluator import Evaluator, zero_based_faces  # noqa: E402


class CudaPath:
    """The three hot-path calls of the baseline on liboryon_b200 (no CPU fallback behind them)."""

    def __init__(self, pointdsc_ckpt: str, device: str = "cuda"):
        from oryon_b200.utils import pcd
        from oryon_b200.utils.pointdsc.init import get_pointdsc_pose, get_pointdsc_solver
        self._pcd, self._pose, self.device = pcd, get_pointdsc_pose, device
        self.solver = get_pointdsc_solver(pointdsc_ckpt, device)

    def match(self, feats_a, feats_q, kp_a, kp_q, threshold, max_corrs, **variant):
        return self._pcd.nn_correspondences_kp(feats_a, feats_q, kp_a, kp_q, threshold, max_corrs, **variant)

    def lift(self, depth, K, xy):
        return self._pcd.lift_pcd(depth, K, xy).cpu()

    def pose(self, pcd_a, pcd_q):
        return self._pose(self.solver, pcd_a, pcd_q, self.device)


def grey(rgb: np.ndarray) -> np.ndarray:
    import cv2 as cv
    return cv.cvtColor(rgb, cv.COLOR_BGR2GRAY)


def sift_keypoints(sift, rgb: np.ndarray):
    """``(kp int16 [n,2] (x, y), descriptors float32 [n,128])`` (sift_nocs.py:85-97)."""
    kp, feats = sift.detectAndCompute(grey(rgb), None)
    feats = np.zeros((0, 128), np.float32) if feats is None else np.asarray(feats)
    return np.asarray([k.pt for k in kp]).reshape(-1, 2).astype(np.int16), feats


def object_mask(item: dict, mask_dir: Optional[str]) -> np.ndarray:
    if mask_dir is not None:                 # masks == 'ours': one PNG per (frame, object), 1 = object
        from PIL import Image
        return np.where(np.asarray(Image.open(os.path.join(mask_dir, item["instance_id"] + ".png"))) == 1, 1, 0)
    return np.where(item["mask"] == item["metadata"]["mask_ids"][0], 1, 0)


def run_baseline(dataset, path, kind: str, mask: str = "oracle", mask_dir: Optional[str] = None, compute_vsd: bool = True,
                 out_file: Optional[str] = None, pose_errors=None, sift=None) -> Evaluator:
    import cv2 as cv
    variant = dict(max_source=1000, keep_empty=True) if kind == "toyl" else {}
    evaluator = Evaluator(f"{kind.upper()} SIFT ({mask})", compute_vsd=compute_vsd, compute_iou=False, pose_errors=pose_errors)
    sift = sift or cv.SIFT_create()
    mask_type = "ovseg" if mask == "ovseg" else "oracle"
    K = torch.tensor(dataset._reader.get_camera()).flatten()
    models, diams, symms = dataset.get_object_info()
    evaluator.add_object_info(zero_based_faces(models), diams, symms)
    evaluator.init_test()
    fp = open(out_file or f"sift_{kind}_{mask}.txt", "w")
    for i in range(len(dataset)):
        inst = dataset.instances[i]
        scene_a, img_a, scene_q, img_q, obj = inst[1], inst[2], inst[3], inst[4], inst[-1]
        instance_id = f"{scene_a}_{img_a}_{scene_q}_{img_q}_{obj}"
        item_a, item_q = dataset.get_item(scene_a, img_a, obj, mask_type), dataset.get_item(scene_q, img_q, obj, mask_type)
        failure = {"cls_id": [obj], "instance_id": [instance_id]}
        solved = False
        if len(item_a["metadata"]["mask_ids"]) > 0 and len(item_q["metadata"]["mask_ids"]) > 0:
            kp_a, feats_a = sift_keypoints(sift, item_a["rgb"])
            kp_q, feats_q = sift_keypoints(sift, item_q["rgb"])
            mask_a, mask_q = object_mask(item_a, mask_dir), object_mask(item_q, mask_dir)
            valid_a, valid_q = mask_a[kp_a[:, 1], kp_a[:, 0]], mask_q[kp_q[:, 1], kp_q[:, 0]]
            if np.count_nonzero(valid_a) > 0 and np.count_nonzero(valid_q) > 0:      # may fail with predicted masks
                corrs = path.match(torch.tensor(feats_a[valid_a == 1]), torch.tensor(feats_q[valid_q == 1]), torch.tensor(kp_a[valid_a == 1]),
                                   torch.tensor(kp_q[valid_q == 1]), 0.25, 500, **variant)
                if corrs.shape[0] > 0:
                    corrs_a, corrs_q = corrs[:, :2].to(torch.long), corrs[:, 2:].to(torch.long)
                    depth_a, depth_q = torch.tensor(np.asarray(item_a["depth"]).astype(np.int32)), torch.tensor(np.asarray(item_q["depth"]).astype(np.int32))
                    pcd_a = path.lift(depth_a.unsqueeze(-1), K, (corrs_a[:, 0], corrs_a[:, 1])) / 1000.    # mm -> m
                    pcd_q = path.lift(depth_q.unsqueeze(-1), K, (corrs_q[:, 0], corrs_q[:, 1])) / 1000.
                    pred_pose = path.pose(pcd_a, pcd_q)
                    gt_a, gt_q = torch.tensor(item_a["metadata"]["poses"][0]), torch.tensor(item_q["metadata"]["poses"][0])
                    pred_q = pred_pose @ gt_a.to(torch.float32)
                    evaluator.register_test({"gt_pose": gt_q.unsqueeze(0), "pred_pose": pred_q.unsqueeze(0), "pred_pose_rel": pred_pose.unsqueeze(0),
                                             "cls_id": [obj], "camera": [K.cpu().numpy()], "depth": [depth_q.cpu().numpy()],
                                             "instance_id": [instance_id]})
                    pose_txt = " ".join([str(n.item()) for n in pred_pose[:3, :].flatten()])
                    fp.write(",".join([item_a["instance_id"], item_q["instance_id"], pose_txt]) + "\n")
                    solved = True
        if not solved:
            print("Problem with pair ", item_a["instance_id"], item_q["instance_id"], " : missing mask")
            evaluator.register_test_failure(failure)
    fp.close()
    return evaluator


def main(argv=None):
    ap = argparse.ArgumentParser(description="SIFT + PointDSC baseline over a dataset pair split")
    ap.add_argument("--dataset-type", required=True, choices=["nocs", "toyl"])
    ap.add_argument("--root", default="data")
    ap.add_argument("--dataset", default=None, help="dataset folder under --root (default: the dataset type)")
    ap.add_argument("--split", default="cross_scene_test")
    ap.add_argument("--obj", default="all")
    ap.add_argument("--mask", default="oracle", help="'oracle' or 'ovseg' (predicted segmentation files of the dataset tree)")
    ap.add_argument("--mask-dir", default=None, help="folder of '<scene> <img> <obj>.png' masks (1 = object), the reference's masks == 'ours'")
    ap.add_argument("--pointdsc", default="pretrained_models/pointdsc", help="folder holding snapshot/PointDSC_3DMatch_release")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--no-vsd", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    from oryon_b200.datasets import NOCSDataset, TOYLDataset
    cfg = dict(dataset=dict(root=args.root, max_corrs=500, img_size=[224, 224],
                            test=dict(name=args.dataset or args.dataset_type, split=args.split, obj=args.obj)),
               test=dict(mask=args.mask, add_description="yes"))
    dataset = {"nocs": NOCSDataset, "toyl": TOYLDataset}[args.dataset_type](cfg, eval=True)
    evaluator = run_baseline(dataset, CudaPath(args.pointdsc, args.device), args.dataset_type, args.mask, args.mask_dir, not args.no_vsd, args.out)
    print(evaluator.test_summary())
    print(evaluator.get_latex_str())


if __name__ == "__main__":
    main()
