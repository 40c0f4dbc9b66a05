"""Readers for the on-disk layout of the reference's TOYL (BOP "toyota light") test data (reference utils/data/toyl.py),
mirroring -- like ``utils/data/nocs.py`` here -- what the test loop and the evaluator consume: per-frame annotations
(``get_part_data``, ``get_item_metadata``), frames (``get_item_data``), object names / models / symmetry sets.

The object models are BOP ``.ply`` files; the reference reads them with the third-party ``plyfile`` package
(utils/data/toyl.py:16, :63), which is not a dependency here: ``read_ply`` parses the two encodings BOP ships (ASCII and
binary little endian) with numpy.  Pinned by ``oracle/make_golden_toyl.py`` -> ``tests/golden/toyl_tree_0.*``.
"""
from __future__ import annotations

import json
import os
from os.path import join
from typing import Dict, Optional, Tuple

import numpy as np
from PIL import Image

from .nocs import get_symmetry_transformations

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
              "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def get_camera() -> np.ndarray:
    """As the reference's ``toyl.get_camera()`` (utils/data/toyl.py:19-21) this returns the NOCS intrinsics; the TOYL
    dataset class does not call it and hard-codes the real ones (datasets.py:573, ``TOYLDataset.K`` here)."""
    return np.asarray([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])


def parse_pair_line(line: str) -> Tuple[str, int, int, int, int, int]:
    """``'test, 1 0, 2 3, 5'`` -> (split, scene_a, img_a, scene_q, img_q, object id) (datasets.py:603-609)."""
    split, idx_a, idx_q, cls_id = line.split(",")
    scene_a, img_a = [int(n) for n in idx_a.split(" ") if n != ""]
    scene_q, img_q = [int(n) for n in idx_q.split(" ") if n != ""]
    return split, scene_a, img_a, scene_q, img_q, int(cls_id)


def read_ply(path: str) -> Dict[str, Dict[str, np.ndarray]]:
    """``{element: {property: array}}`` of a PLY file (ASCII or binary little endian); list properties (faces) come back as
    an ``[n, k]`` integer array when every list has the same length, else as an object array of arrays."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                elements[-1][2].append((tok[-1], tok[1:-1]))      # (name, ['float'] or ['list', count type, item type])
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian"):
            raise ValueError(f"{path}: PLY format {fmt!r} not supported")
        out: Dict[str, Dict[str, np.ndarray]] = {}
        for name, count, props in elements:
            scalar = all(len(t) == 1 for _, t in props)
            if fmt == "binary_little_endian" and scalar:
                rec = np.frombuffer(f.read(count * np.dtype([(n, "<" + _PLY_TYPES[t[0]]) for n, t in props]).itemsize),
                                    dtype=[(n, "<" + _PLY_TYPES[t[0]]) for n, t in props], count=count)
                out[name] = {n: np.array(rec[n]) for n, _ in props}
                continue
            cols = {n: [] for n, _ in props}
            for _ in range(count):
                toks = f.readline().split() if fmt == "ascii" else None
                pos = 0
                for n, t in props:
                    if len(t) == 1:
                        if fmt == "ascii":
                            v = np.dtype(_PLY_TYPES[t[0]]).type(toks[pos])
                            pos += 1
                        else:
                            v = np.frombuffer(f.read(np.dtype(_PLY_TYPES[t[0]]).itemsize), "<" + _PLY_TYPES[t[0]])[0]
                    else:
                        ct, it = _PLY_TYPES[t[1]], _PLY_TYPES[t[2]]
                        if fmt == "ascii":
                            k = int(toks[pos])
                            v = np.array(toks[pos + 1:pos + 1 + k], dtype=it)
                            pos += 1 + k
                        else:
                            k = int(np.frombuffer(f.read(np.dtype(ct).itemsize), "<" + ct)[0])
                            v = np.frombuffer(f.read(k * np.dtype(it).itemsize), "<" + it).copy()
                    cols[n].append(v)
            out[name] = {}
            for n, t in props:
                if len(t) == 1:
                    out[name][n] = np.array(cols[n], dtype=_PLY_TYPES[t[0]])
                elif len({len(v) for v in cols[n]}) <= 1:
                    out[name][n] = np.stack(cols[n], axis=0) if cols[n] else np.zeros((0, 0), _PLY_TYPES[t[2]])
                else:
                    arr = np.empty(len(cols[n]), dtype=object)
                    arr[:] = cols[n]
                    out[name][n] = arr
        return out


def get_obj_rendering(root: str, obj_id: int) -> dict:
    """Object model for the evaluator / the rasteriser (utils/data/toyl.py:52-82): ``pts`` (already mm), ``normals``, 0-based
    ``faces`` as in the PLY file."""
    ply = read_ply(os.path.join(root, "models_bop", "obj_{:06d}.ply".format(obj_id)))
    v, face = ply["vertex"], ply["face"]
    faces = face["vertex_indices"] if "vertex_indices" in face else face["vertex_index"]
    return {"pts": np.stack((v["x"], v["y"], v["z"]), axis=1), "normals": np.stack((v["nx"], v["ny"], v["nz"]), axis=1),
            "faces": np.stack([f for f in faces], axis=0)}


def get_obj_names(root: str) -> dict:
    with open(join(root, "models_name.json")) as f:
        return json.load(f)


def get_part_data(root: str) -> dict:
    """``{'<scene>_<img>': {'<object id>': {pose (translation in metres), cls_id, box [x,y,w,h], mask_idx}}}`` from the BOP
    ``scene_gt.json`` / ``scene_gt_info.json`` of every scene (utils/data/toyl.py:93-137).  ``mask_idx`` is the 1-based
    position of the object in the frame's annotation list -- its label in the merged ``mask_visib`` image.  As in the
    reference, the frame entry is created at the first object and a second instance of an object id overwrites the first."""
    new_data = {}
    for scene_folder in os.listdir(join(root, "split", "test")):
        with open(join(root, "split", "test", scene_folder, "scene_gt.json")) as fa, \
                open(join(root, "split", "test", scene_folder, "scene_gt_info.json")) as fm:
            data, meta = json.load(fa), json.load(fm)
        for img_k, img_data in data.items():
            for i, (obj, obj_meta) in enumerate(zip(img_data, meta[img_k])):
                pose = np.eye(4)
                pose[:3, :3] = np.asarray(obj["cam_R_m2c"]).reshape(3, 3)
                pose[:3, 3] = np.asarray(obj["cam_t_m2c"]) / 1000.
                cls_id = int(obj["obj_id"])
                img_dk = f"{int(scene_folder)}_{int(img_k)}"
                if i == 0:
                    new_data[img_dk] = {}
                new_data[img_dk][f"{int(cls_id)}"] = {"pose": pose, "cls_id": cls_id, "box": obj_meta["bbox_visib"], "mask_idx": i + 1}
    return new_data


def get_item_metadata(root: str, scene_id: int, img_id: int, pose_annots: dict, cls_names_dict: dict, cls_id: Optional[int] = None) -> dict:
    img_annots = pose_annots[f"{scene_id}_{img_id}"]
    cls_ids, mask_ids, cls_names, cls_descs, poses, boxes = [], [], [], [], [], []
    for obj_id in list(img_annots.keys()):
        if cls_id is not None and int(obj_id) != int(cls_id):
            continue
        cls_ids.append(int(obj_id))
        mask_ids.append(img_annots[obj_id]["mask_idx"])
        cls_names.append(cls_names_dict[obj_id][0])
        cls_descs.append(cls_names_dict[obj_id][1:])
        poses.append(img_annots[obj_id]["pose"])
        boxes.append(img_annots[obj_id]["box"])
    return {"cls_ids": cls_ids, "mask_ids": mask_ids, "cls_names": cls_names, "cls_descs": cls_descs, "poses": poses, "boxes": boxes}


def get_item_data(root: str, scene_id: int, img_id: int, pose_annots: dict, cls_names: dict, cls_id: Optional[int] = None,
                  mask_type: Optional[str] = "oracle", hf_depth: bool = False) -> dict:
    """One decoded frame (utils/data/toyl.py:165-213): ``rgb`` uint8 HWC, ``mask`` label image, ``depth`` (mm)."""
    metadata = get_item_metadata(root, scene_id, img_id, pose_annots, cls_names, cls_id=cls_id)
    base = join(root, "split", "test", f"{scene_id:06d}")
    img = np.asarray(Image.open(join(base, "rgb", f"{img_id:06d}.png")).convert("RGB"))
    if mask_type == "oracle":
        mask = np.asarray(Image.open(join(base, "mask_visib", f"{img_id:06d}.png")).convert("L"))
    elif mask_type == "ovseg":
        mask = np.asarray(Image.open(join(base, "mask_pred", f"{img_id:06d}.png")).convert("L"))
    elif mask_type in ("san", "oryon"):
        folder = "san_name" if mask_type == "san" else "oryon"
        mask = np.asarray(Image.open(join(root, folder, f"{scene_id} {img_id} {cls_id}.png")).convert("L"))
        mask = np.where(mask == 1, metadata["mask_ids"][0], 255)
    else:
        raise RuntimeError(f"Mask type {mask_type} not implemented.")
    depth = np.asarray(Image.open(join(base, "hf_depth" if hf_depth else "depth", f"{img_id:06d}.png")))
    return {"rgb": img, "mask": mask, "depth": depth, "metadata": metadata, "instance_id": f"{scene_id} {img_id} {cls_id}"}


def get_obj_data(root: str) -> Tuple[dict, dict, dict]:
    """``(models, diameters [mm], symmetry sets)`` keyed by integer object id, for every ``.ply`` under ``models_bop``
    (utils/data/toyl.py:215-236; symmetries discretised with ``max_sym_disc_step=0.05``)."""
    with open(join(root, "models_bop", "models_info.json")) as f:
        models_info = json.load(f)
    models, diams, symms = {}, {}, {}
    for obj_file in [f for f in os.listdir(join(root, "models_bop")) if ".ply" in f]:
        obj_id = int(os.path.splitext(obj_file[4:])[0])
        info = models_info[str(obj_id)]
        models[obj_id] = get_obj_rendering(root, obj_id)
        diams[obj_id] = info["diameter"]
        symms[obj_id] = get_symmetry_transformations(info, max_sym_disc_step=0.05)
    return models, diams, symms
