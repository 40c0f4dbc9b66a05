"""Mirror of the reference ``utils/evaluator.py`` for the test path: ``Evaluator`` with the same method
names, metric keys, failure bookkeeping and LaTeX / JSON output; ``dict_from_preds`` of
scripts/evaluation/compute_metrics.py:14-49 (the CSV wire format ``FPM_Pipeline.add_pred_pose`` writes).

The pose-error arithmetic of ``register_eval`` (utils/evaluator.py:206-288: R/T error, ADD or ADD-S, MSSD, MSPD) runs in
liboryon_b200.so (``oryon_eval_pose_errors``, csrc/eval.cu) for a whole batch of pairs at once; what stays here is the
bookkeeping (lists, thresholds, means).  With ``compute_vsd=True`` the VSD errors (bop_toolkit_lib/pose_error.py:17-96) come
from ``oryon_eval_vsd``: a CUDA z-buffer rasteriser in place of the reference's OpenGL renderer
(bop_toolkit_lib/renderer_vispy.py) followed by the reference's visibility / cost arithmetic; object models then need
``faces``.  AR = (MSSD + MSPD + VSD) / 3 as in utils/evaluator.py:286.
"""
from __future__ import annotations

import ctypes
import json
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

POSE_RECALL_TH = [(5, 10), (10, 20), (15, 30)]


def format_sym_set(syms) -> np.ndarray:
    """BOP symmetry list ``[{'R': [3,3], 't': [3,1]}]`` -> ``[S,3,4]`` (bop_toolkit_lib/misc.py:402-411)."""
    return np.stack([np.concatenate([np.asarray(s["R"], dtype=np.float64).reshape(3, 3),
                                     np.asarray(s["t"], dtype=np.float64).reshape(3, 1)], axis=1) for s in syms], axis=0)


def get_diameter(pcd: np.ndarray) -> float:
    """Largest bounding-box side (utils/pcd.py:16-20): the 'diameter' ADD(S)-0.1d is measured against."""
    xyz = np.asarray(pcd)[:, :3]
    return max(np.max(xyz, axis=0) - np.min(xyz, axis=0))


def dict_from_preds(perf_file: str):
    """Reads the prediction CSV: ``id_a,id_q,<12 floats>[,iou_a,iou_q]`` per line with ids ``'scene img obj'``
    (scripts/evaluation/compute_metrics.py:14-49) -> ``(preds {key: [3,4]}, ious_a, ious_q, iou_present)``."""
    preds, ious_a, ious_q, iou_present = {}, {}, {}, True
    with open(perf_file, "r") as fh:
        for line in fh.readlines():
            tokens = line.split(",")
            if len(tokens) == 3:
                (id_a, id_q, pose_line), iou_present = tokens, False
            elif len(tokens) == 5:
                id_a, id_q, pose_line, iou_a, iou_q = tokens
                iou_present = True
            else:
                raise RuntimeError(" Anomaly in line: " + line)
            scene_a, img_a, obj_a = id_a.split(" ")
            scene_q, img_q, _ = id_q.split(" ")
            key = "{}_{}_{}_{}_{}".format(scene_a, img_a, scene_q, img_q, obj_a)
            preds[key] = np.asarray([float(p) for p in pose_line.split(" ")]).reshape(3, 4)
            if iou_present:
                ious_a[key], ious_q[key] = float(iou_a), float(iou_q)
    return preds, ious_a, ious_q, iou_present


def zero_based_faces(obj_models: dict) -> dict:
    """Face indices as the rasteriser needs them: OBJ files (NOCS) count from 1, PLY files (TOYL) from 0.  The NOCS reader marks its
    models with ``faces_base``; unmarked models are shifted only when their indices cannot be 0-based (minimum >= 1 and the maximum
    equal to the vertex count)."""
    out = {}
    for k, m in obj_models.items():
        m = dict(m)
        if "faces" in m and len(m["faces"]):
            base = m.pop("faces_base", None)
            if base is None:
                base = 1 if (int(np.min(m["faces"])) >= 1 and int(np.max(m["faces"])) == len(m["pts"])) else 0
            if base:
                m["faces"] = np.asarray(m["faces"]) - int(base)
        out[k] = m
    return out


class CudaPoseErrors:
    """Object models / symmetry sets resident on the GPU + ``oryon_eval_pose_errors``.  ``__call__(cls_ids, pred [P,4,4],
    gt [P,4,4], cams [P,3,3]) -> float64 [P,6]``: R error (deg), T error (cm), ADD or ADD-S (m), ADD-S flag, MSSD (mm),
    MSPD (px)."""

    def __init__(self, device=None):
        from .._torch_glue import require_cuda
        self.device = torch.device(device) if device is not None else require_cuda()
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._ids: Dict[object, int] = {}

    def add_object(self, key, pts_mm: np.ndarray, syms: np.ndarray) -> None:
        from .. import _lib
        pts = np.ascontiguousarray(np.asarray(pts_mm, dtype=np.float64)[:, :3])
        sym = np.ascontiguousarray(np.asarray(syms, dtype=np.float64).reshape(-1, 12))
        oid = self._ids.setdefault(key, len(self._ids))
        dbl = ctypes.POINTER(ctypes.c_double)
        _lib.check(_lib.load().oryon_eval_set_object(_lib.handle(self.device.index), oid, pts.ctypes.data_as(dbl), pts.shape[0],
                                                     sym.ctypes.data_as(dbl), sym.shape[0]))

    def add_mesh(self, key, faces: np.ndarray) -> None:
        from .. import _lib
        f = np.ascontiguousarray(np.asarray(faces, dtype=np.int32).reshape(-1, 3))
        _lib.check(_lib.load().oryon_eval_set_object_mesh(_lib.handle(self.device.index), self._ids[key],
                                                          f.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), f.shape[0]))

    def vsd(self, cls_ids: Sequence, pred: np.ndarray, gt: np.ndarray, cams: np.ndarray, depths: Sequence, diameters: Sequence[float],
            delta: float, taus: Sequence[float]) -> np.ndarray:
        """``[P, len(taus)]`` VSD errors; ``depths``: P test depth images ``[H,W]`` in mm (integer or float, 0 = missing)."""
        from .. import _lib
        from .._torch_glue import ptr, stream_ptr
        P, dev = len(cls_ids), self.device
        ids = (ctypes.c_int32 * P)(*[self._ids[c] for c in cls_ids])
        d = np.stack([np.asarray(x).squeeze() for x in depths])
        is_f32 = not np.issubdtype(d.dtype, np.integer)
        td = torch.as_tensor(np.ascontiguousarray(d, dtype=np.float32 if is_f32 else np.int32)).to(dev)
        tp = torch.as_tensor(np.ascontiguousarray(pred, dtype=np.float64).reshape(P, 16)).to(dev)
        tg = torch.as_tensor(np.ascontiguousarray(gt, dtype=np.float64).reshape(P, 16)).to(dev)
        tk = torch.as_tensor(np.ascontiguousarray(cams, dtype=np.float64).reshape(P, 9)).to(dev)
        out = torch.empty(P, len(taus), dtype=torch.float64, device=dev)
        c_taus = (ctypes.c_double * len(taus))(*[float(t) for t in taus])
        c_diam = (ctypes.c_double * P)(*[float(x) for x in diameters])
        _lib.check(_lib.load().oryon_eval_vsd(_lib.handle(dev.index), P, ids, ptr(tp), ptr(tg), ptr(tk), ptr(td), int(is_f32), d.shape[1],
                                              d.shape[2], float(delta), c_taus, len(taus), c_diam, ptr(out), stream_ptr(dev)))
        return out.cpu().numpy()

    def __call__(self, cls_ids: Sequence, pred: np.ndarray, gt: np.ndarray, cams: np.ndarray) -> np.ndarray:
        from .. import _lib
        from .._torch_glue import ptr, stream_ptr
        P = len(cls_ids)
        ids = (ctypes.c_int32 * P)(*[self._ids[c] for c in cls_ids])
        dev = self.device
        tp = torch.as_tensor(np.ascontiguousarray(pred, dtype=np.float64).reshape(P, 16)).to(dev)
        tg = torch.as_tensor(np.ascontiguousarray(gt, dtype=np.float64).reshape(P, 16)).to(dev)
        tk = torch.as_tensor(np.ascontiguousarray(cams, dtype=np.float64).reshape(P, 9)).to(dev)
        out = torch.empty(P, 6, dtype=torch.float64, device=dev)
        _lib.check(_lib.load().oryon_eval_pose_errors(_lib.handle(dev.index), P, ids, ptr(tp), ptr(tg), ptr(tk), ptr(out), stream_ptr(dev)))
        return out.cpu().numpy()


class Evaluator(object):
    """Helper class used to evaluate pose metrics (utils/evaluator.py:84-432)."""

    def __init__(self, exp_tag: str, compute_vsd: bool = True, compute_iou: bool = True, *, pose_errors: Optional[Callable] = None,
                 device=None):
        self.exp_tag = exp_tag
        self.mssd_rec = np.arange(0.05, 0.51, 0.05)
        self.mspd_rec = np.arange(5, 51, 5)
        self.compute_vsd, self.compute_iou = compute_vsd, compute_iou
        if self.compute_vsd:   # utils/evaluator.py:96-101
            self.vsd_taus = list(np.arange(0.05, 0.51, 0.05))
            self.vsd_rec = np.arange(0.05, 0.51, 0.05)
            self.vsd_delta = 15.
        self.pose_recall_th = POSE_RECALL_TH
        self.metrics: Dict[str, list] = {}
        self.counts: Dict[str, list] = {}
        self._pose_errors = pose_errors            # injected checker in CPU tests; the CUDA library otherwise
        self._device = device

    # ---- object info ------------------------------------------------------------------------------------
    def add_object_info(self, obj_models: dict, obj_diams: dict, obj_symms: dict):
        """Models / diameters in mm, BOP symmetry lists (utils/evaluator.py:111-119).  Meshes marked (or recognisable) as 1-based OBJ
        indices are shifted for the rasteriser (``zero_based_faces``); the reference hands them to OpenGL as they are."""
        obj_models = zero_based_faces(obj_models)
        self.obj_models, self.obj_diams = obj_models, obj_diams
        self.obj_symms = {k: format_sym_set(s) for k, s in obj_symms.items()}
        self.add_diams = {k: get_diameter(m["pts"]) / 1000. for k, m in obj_models.items()}
        if self._pose_errors is None:
            self._pose_errors = CudaPoseErrors(self._device)
        if hasattr(self._pose_errors, "add_object"):
            for k, m in obj_models.items():
                self._pose_errors.add_object(k, m["pts"], self.obj_symms[k])
                if self.compute_vsd:
                    if "faces" not in m:
                        raise ValueError(f"compute_vsd=True needs a triangle mesh: object {k} has no 'faces'")
                    self._pose_errors.add_mesh(k, m["faces"])

    def get_obj_info(self, obj_id):
        return self.obj_models[obj_id], self.obj_diams[obj_id], self.obj_symms[obj_id]

    # ---- storage ----------------------------------------------------------------------------------------
    def clear(self):
        self.metrics, self.counts = {}, {}

    def init_training(self):
        self.clear()
        if self.compute_iou:
            for k in ("Anchor IoU", "Query IoU", "Mean IoU", "IoU > .25", "IoU > .5", "IoU > .75"):
                self.metrics[k] = []

    def init_validation(self):
        self.init_training()
        for k in ("R error", "T error", "ADD(S)-0.1d") + (("AR", "VSD") if self.compute_vsd else ()) + ("MSSD", "MSPD"):
            self.metrics[k] = []
        for k in ("Missing segm", "Failed pose", "Zero pose"):
            self.counts[k] = []
        for r_th, t_th in self.pose_recall_th:
            self.metrics[f"Recall ({r_th}deg, {t_th}cm)"] = []

    def init_test(self):
        self.init_validation()
        self.metrics["instance_id"] = []
        self.metrics["cls_id"] = []

    # ---- registration -----------------------------------------------------------------------------------
    @staticmethod
    def _np(t) -> np.ndarray:
        return t.clone().detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    def register_train(self, results: dict, clear: bool = False):
        if clear:
            self.clear()
            self.init_training()
        if self.compute_iou:
            iou_a, iou_q = self._np(results["iou_a"]), self._np(results["iou_q"])
            mean_iou = self._np((results["iou_a"] + results["iou_q"]) / 2.)
            self.metrics["Anchor IoU"].extend(iou_a.tolist())
            self.metrics["Query IoU"].extend(iou_q.tolist())
            self.metrics["Mean IoU"].extend(mean_iou.tolist())
            self.metrics["IoU > .25"].extend((mean_iou > 0.25).astype(int).tolist())
            self.metrics["IoU > .5"].extend((mean_iou > 0.5).astype(int).tolist())
            self.metrics["IoU > .75"].extend((mean_iou > 0.75).astype(int).tolist())

    def register_eval(self, results: dict, clear: bool = False):
        """Any number of pairs per call (the reference's ``test_step`` registers one at a time, :321-331; a batch gives
        the same lists)."""
        self.register_train(results, clear)
        pred_poses = self._np(results["pred_pose"]).copy()
        gt_poses = self._np(results["gt_pose"])
        pred_poses_rel = self._np(results["pred_pose_rel"])
        for idx, rel in enumerate(pred_poses_rel):
            self.counts["Missing segm"].append(0)
            zero_pose = int(np.count_nonzero(rel) <= 1)
            self.counts["Failed pose"].append(int((rel == np.eye(4)).all()))
            self.counts["Zero pose"].append(zero_pose)
            if zero_pose == 1:
                pred_poses[idx] = np.eye(4)
        cams = np.stack([np.asarray(c, dtype=np.float64).reshape(3, 3) for c in results["camera"]])
        err = self._pose_errors(list(results["cls_id"]), pred_poses, gt_poses, cams)
        err_R, err_T = err[:, 0], err[:, 1]
        self.metrics["R error"].extend(err_R.tolist())
        self.metrics["T error"].extend(err_T.tolist())
        for r_th, t_th in self.pose_recall_th:
            ok = np.logical_and(err_R <= r_th, err_T <= t_th).astype(float)
            self.metrics[f"Recall ({r_th}deg, {t_th}cm)"].extend(ok.tolist())
        vsd_errs = None
        if self.compute_vsd:
            vsd_errs = self._pose_errors.vsd(list(results["cls_id"]), pred_poses, gt_poses, cams, results["depth"],
                                             [self.obj_diams[c] for c in results["cls_id"]], self.vsd_delta, self.vsd_taus)
        for i, cls_id in enumerate(results["cls_id"]):
            self.metrics["ADD(S)-0.1d"].append(float(err[i, 2] <= self.add_diams[cls_id] * 0.1))
            mean_mssd = (err[i, 4] < self.mssd_rec * self.obj_diams[cls_id]).mean()
            mean_mspd = (err[i, 5] < self.mspd_rec).mean()
            self.metrics["MSSD"].append(mean_mssd)
            self.metrics["MSPD"].append(mean_mspd)
            if self.compute_vsd:   # utils/evaluator.py:279-286: every (tau, recall threshold) pair counts
                mean_vsd = np.stack([vsd_errs[i] < rec_i for rec_i in self.vsd_rec], axis=1).mean()
                self.metrics["VSD"].append(mean_vsd)
                self.metrics["AR"].append((mean_mssd + mean_mspd + mean_vsd) / 3.)

    def register_test(self, results: dict, clear: bool = False):
        self.register_eval(results, clear)
        self.metrics["cls_id"].extend(results["cls_id"])
        self.metrics["instance_id"].extend(results["instance_id"])

    def register_valid_failure(self, results):
        for k in ("R error", "T error", "ADD(S)-0.1d") + (("VSD", "AR") if self.compute_vsd else ()) + ("MSSD", "MSPD"):
            self.metrics[k].append(0.)
        if self.compute_iou:
            self.metrics["Anchor IoU"].extend(self._np(results["iou_a"]).tolist())
            self.metrics["Query IoU"].extend(self._np(results["iou_q"]).tolist())
            for k in ("Mean IoU", "IoU > .25", "IoU > .5", "IoU > .75"):
                self.metrics[k].append(0.)
        self.counts["Missing segm"].append(1)
        self.counts["Failed pose"].append(0)
        self.counts["Zero pose"].append(0)
        for r_th, t_th in self.pose_recall_th:
            self.metrics[f"Recall ({r_th}deg, {t_th}cm)"].extend([0])

    def register_test_failure(self, results: dict):
        self.register_valid_failure(results)
        self.metrics["cls_id"].extend(results["cls_id"])
        self.metrics["instance_id"].extend(results["instance_id"])

    # ---- summaries --------------------------------------------------------------------------------------
    def save(self, file):
        all_dict = dict()
        all_dict.update(self.metrics)
        all_dict.update(self.counts)
        json.dump(all_dict, file)

    def get_means(self):
        return {name: np.asarray(value).mean() for name, value in self.metrics.items()
                if name not in ["cls_id", "instance_id"] and len(value) > 0}

    get_log_means = get_means

    def get_obj_means(self, cls_id):
        idxs = np.asarray(self.metrics["cls_id"]) == cls_id
        return {name: np.asarray(value)[idxs].mean() for name, value in self.metrics.items()
                if name not in ["cls_id", "instance_id"] and len(value) > 0}

    def _row(self, tag, means) -> str:
        head = f"{means['AR']*100:.1f} & {means['VSD']*100:.1f}" if self.compute_vsd else "- & -"
        s = f"{tag} & {head} & {means['MSSD']*100:.1f} & {means['MSPD']*100:.1f} & {means['ADD(S)-0.1d']*100:.1f} &"
        return s + (f" {means['Mean IoU']*100:.1f} \\\\" if self.compute_iou else " - \\\\")

    def test_summary(self):
        for cls_id in np.unique(self.metrics["cls_id"]).tolist():
            print(self._row(cls_id, self.get_obj_means(cls_id)))

    def get_latex_str(self):
        return self._row(self.exp_tag, self.get_means()) + " \n"
