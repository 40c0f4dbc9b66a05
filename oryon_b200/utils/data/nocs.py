"""Readers for the on-disk layout of the reference's NOCS (REAL275) test data (reference utils/data/nocs.py): the input
side of the inference path, so that ``run_test.py`` / ``FPM_Pipeline`` can be pointed at a mounted dataset.

Only what the test loop and the evaluator consume is mirrored -- frame readers (``get_item_data``, ``get_item_metadata``),
per-frame pose annotations (``get_part_data``), object names / models / symmetries (``get_obj_names``,
``get_obj_rendering``, ``get_obj_data``) -- with the reference's return structures, so that ``GpuCollate`` and the
``Evaluator`` mirror take them unchanged.  Point-cloud helpers used by the data-preparation scripts are not.
Pinned by ``oracle/make_golden_nocs.py`` (the reference's own readers on a synthetic tree) -> ``tests/golden/nocs_tree_0.*``.
"""
from __future__ import annotations

import json
import math
import os
import pickle
from os.path import join
from typing import Dict, List, Optional, Tuple

import numpy as np
from PIL import Image


def get_camera() -> np.ndarray:
    """Intrinsics of the real NOCS frames (utils/data/nocs.py:17-18, datasets.py:398)."""
    return np.asarray([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])


def parse_pair_line(line: str) -> Tuple[str, int, int, int, int, int, str]:
    """``'real_test, 1 0, 2 3, 6 mug_name'`` -> (split, scene_a, img_a, scene_q, img_q, category id, object name)
    (datasets.py:430-437)."""
    split, idx_a, idx_q, cat = line.split(",")
    cat_id, obj_name = cat.strip().split(" ")
    scene_a, img_a = [int(n) for n in idx_a.split(" ") if n != ""]
    scene_q, img_q = [int(n) for n in idx_q.split(" ") if n != ""]
    return split, scene_a, img_a, scene_q, img_q, int(cat_id), obj_name


def get_obj_names(root: str) -> dict:
    with open(join(root, "obj_names.json")) as f:
        return json.load(f)


def get_part_data(root: str) -> Dict[str, np.ndarray]:
    """All per-frame object poses ``{'<scene>_<img>': [n_obj,4,4]}`` (utils/data/nocs.py:91-106)."""
    poses = {}
    for img_file in os.listdir(join(root, "gts", "real_test")):
        with open(join(root, "gts", "real_test", img_file), "rb") as f:
            data = pickle.load(f)["gt_RTs"]
        scene_id, img_id = os.path.splitext(img_file)[0].split("_")[-2:]
        poses[f"{int(scene_id)}_{int(img_id)}"] = data
    return poses


def get_item_metadata(root: str, scene_id: int, img_id: int, pose_annots: dict, obj_names: dict, obj_name: Optional[str] = None) -> dict:
    """Per-frame annotations (utils/data/nocs.py:178-227): the NOCS poses with their scale divided out ROW-wise as the
    reference does (``R / norm(R, axis=1)`` broadcasts over columns), class ids / names / descriptions, mask ids and
    detection boxes -- of every object of the frame, or of ``obj_name`` only."""
    poses = []
    for pose in pose_annots[f"{scene_id}_{img_id}"]:
        new_pose = pose.copy()
        new_pose[:3, :3] = new_pose[:3, :3] / np.linalg.norm(new_pose[:3, :3], axis=1)
        poses.append(new_pose)
    cls_ids, mask_ids, cls_names, cls_descs, dets = [], [], [], [], []
    stem = join(root, "split/real_test", f"scene_{scene_id}/{img_id:04d}")
    with open(stem + "_meta.txt") as fm, open(stem + "_detection.txt") as fd:
        for i, (meta_line, det_line) in enumerate(zip(fm.readlines(), fd.readlines())):
            mask_id, cls_id, cur = meta_line.split(" ")
            cur = cur.strip()
            if obj_name is not None:
                if cur != obj_name:
                    continue
                poses = [poses[i]]
            cls_ids.append(int(cls_id))
            mask_ids.append(int(mask_id))
            cls_names.append(obj_names[cur][0])
            cls_descs.append(obj_names[cur][1:])
            x, y, w, h = [int(v) for v in det_line.split(" ")[1:]]
            dets.append((x, y, w, h))
    return {"cls_ids": cls_ids, "mask_ids": mask_ids, "cls_names": cls_names, "cls_descs": cls_descs, "poses": poses, "boxes": dets}


def get_item_data(root: str, scene_id: int, img_id: int, pose_annots: dict, obj_names: dict, obj_name: Optional[str] = None,
                  mask_type: Optional[str] = "oracle", hf_depth: bool = False) -> dict:
    """One decoded frame (utils/data/nocs.py:229-278): ``rgb`` uint8 HWC, ``mask`` label image (NOCS convention: object
    label, 255 elsewhere; the 'san' / 'oryon' prediction files hold 1 for the object and are converted), ``depth`` (mm)."""
    metadata = get_item_metadata(root, scene_id, img_id, pose_annots, obj_names, obj_name)
    base = join(root, "split/real_test", f"scene_{scene_id}/{img_id:04d}")
    img = np.asarray(Image.open(base + "_color.png").convert("RGB"))
    if mask_type == "oracle":
        mask = np.asarray(Image.open(base + "_mask.png").convert("L"))
    elif mask_type == "ovseg":
        mask = np.asarray(Image.open(base + "_pred_mask.png").convert("L"))
    elif mask_type in ("san", "oryon"):
        folder = "san_name" if mask_type == "san" else "oryon"
        mask = np.asarray(Image.open(join(root, folder, f"{scene_id} {img_id} {obj_name}.png")).convert("L"))
        mask = np.where(mask == 1, metadata["mask_ids"][0], 255)
    else:
        raise RuntimeError(f"Mask type {mask_type} not implemented.")
    depth = np.asarray(Image.open(base + ("_hfdepth.png" if hf_depth else "_depth.png")))
    return {"rgb": img, "mask": mask, "depth": depth, "metadata": metadata, "instance_id": f"{scene_id} {img_id} {obj_name}"}


def get_obj_rendering(root: str, obj_id: str) -> dict:
    """Object model for the evaluator / the rasteriser (utils/data/nocs.py:59-89): ``pts`` in mm, ``normals``, 1-based
    ``faces`` as they stand in the OBJ file."""
    base = join(root, "obj_models", "real_test", obj_id)
    with open(base + "_vertices.txt") as f:
        pts = [[float(t) for t in line.split(" ")[:3]] for line in f.readlines()]
    with open(base + "_normals.txt") as f:
        normals = [[float(t) for t in line.split(" ")[:3]] for line in f.readlines()]
    with open(base + ".obj") as f:
        faces = [[int(tok.split("/")[0]) for tok in line.split(" ")[1:4]] for line in f.readlines() if line.startswith("f")]
    # "faces_base": not in the reference's dict -- tells the rasteriser's caller that these indices count from 1 (OBJ convention)
    return {"pts": np.asarray(pts) * 1000, "normals": np.asarray(normals), "faces": np.asarray(faces), "faces_base": 1}


def _rotation_matrix(angle: float, direction) -> np.ndarray:
    """3x3 rotation about ``direction`` in the operation order of bop_toolkit_lib/transform.py ``rotation_matrix``."""
    sina, cosa = math.sin(angle), math.cos(angle)
    d = np.array(direction[:3], dtype=np.float64, copy=True)
    d /= math.sqrt(np.dot(d, d))
    R = np.diag([cosa, cosa, cosa])
    R += np.outer(d, d) * (1.0 - cosa)
    d *= sina
    R += np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    return R


def get_symmetry_transformations(model_info: dict, max_sym_disc_step: float = 0.01) -> List[dict]:
    """BOP symmetry set of an object (bop_toolkit_lib/misc.py:43-90): identity + discrete symmetries, each combined with
    the discretised continuous ones (``ceil(pi / max_sym_disc_step)`` steps about the axis)."""
    trans_disc = [{"R": np.eye(3), "t": np.array([[0, 0, 0]]).T}]
    for sym in model_info.get("symmetries_discrete", []):
        s = np.reshape(sym, (4, 4))
        trans_disc.append({"R": s[:3, :3], "t": s[:3, 3].reshape((3, 1))})
    trans_cont = []
    for sym in model_info.get("symmetries_continuous", []):
        axis = np.array(sym["axis"])
        offset = np.array(sym["offset"]).reshape((3, 1))
        steps = int(np.ceil(np.pi / max_sym_disc_step))
        step = 2.0 * np.pi / steps
        for i in range(steps):
            R = _rotation_matrix(i * step, axis)
            trans_cont.append({"R": R, "t": -R.dot(offset) + offset})
    trans = []
    for td in trans_disc:
        if trans_cont:
            for tc in trans_cont:
                trans.append({"R": tc["R"].dot(td["R"]), "t": tc["R"].dot(td["t"]) + tc["t"]})
        else:
            trans.append(td)
    return trans


def get_obj_data(root: str) -> Tuple[dict, dict, dict]:
    """``(models, diameters [mm], symmetry sets)`` of every object in ``models_info.json`` (utils/data/nocs.py:125-141;
    symmetries discretised with ``max_sym_disc_step=0.05`` as there)."""
    with open(join(root, "obj_models", "real_test", "models_info.json")) as f:
        models_info = json.load(f)
    models, diams, symms = {}, {}, {}
    for name, info in models_info.items():
        models[name] = get_obj_rendering(root, name)
        diams[name] = info["diameter"]
        symms[name] = get_symmetry_transformations(info, max_sym_disc_step=0.05)
    return models, diams, symms
