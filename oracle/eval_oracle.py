"""CPU restatement (numpy) of the pose-error arithmetic behind ``Evaluator.register_eval`` (reference
utils/evaluator.py:206-288).  TEST INFRASTRUCTURE: only tests/ import this; it is the checker for the CUDA
``oryon_eval_pose_errors`` and is itself pinned to reference outputs (tests/golden/eval_0.npz, written by
oracle/make_golden_eval.py from the unmodified reference functions).

Reference quirks restated on purpose (they decide the published numbers):
* ADD / ADD-S transform the model in float16 (utils/pcd.py:127-133 ``np_transform_pcd``); ADD also takes the norm and the
  mean in float16 (utils/metrics.py:205-218), ADD-S measures float64 nearest-neighbour distances between the float16
  clouds (:220-234).
* MSSD / MSPD round the poses to float16 and the translation to float16 millimetres (utils/evaluator.py:258-261), then
  work in float64 -- on the FIRST THREE model points only: ``np_transform`` slices ``pts[:, :3]`` of a ``[1,N,3]``
  array (bop_toolkit_lib/pose_error.py:337-351, called from :391/:396/:364).
"""
import numpy as np


def rt_errors(pred: np.ndarray, gt: np.ndarray):
    """Rotation error (degrees) and translation error (cm) of one pose pair, utils/metrics.py:236-259."""
    # dtypes are kept: test_step hands a float32 prediction and a float64 ground truth (pipeline.py:320-326), so the
    # reference normalises R1 in float32 -- visible as ~1e-7 / sin(theta) rad of noise in the angle
    p, g = np.asarray(pred), np.asarray(gt)
    R1 = p[:3, :3] / np.cbrt(np.linalg.det(p[:3, :3]))
    R2 = g[:3, :3] / np.cbrt(np.linalg.det(g[:3, :3]))
    arg = np.clip((np.trace(R1 @ R2.T) - 1) / 2, -1 + 1e-12, 1 - 1e-12)
    theta = np.arccos(arg) * 180 / np.pi
    if np.isnan(theta):
        theta = 180.0
    return float(theta), float(np.linalg.norm(p[:3, 3] - g[:3, 3]) * 100)


def _transform16(pts_m: np.ndarray, pose: np.ndarray) -> np.ndarray:
    """utils/pcd.py:127-133 in explicit arithmetic: operands rounded to float16, each coordinate accumulated in float32
    over k = 0,1,2 (numpy's half dot product), rounded to float16, translation added in float16."""
    p = pts_m.astype(np.float16).astype(np.float32)
    r = pose[:3, :3].astype(np.float16).astype(np.float32)
    t = pose[:3, 3].astype(np.float16).astype(np.float32)
    out = np.empty((p.shape[0], 3), np.float16)
    for j in range(3):
        acc = np.zeros(p.shape[0], np.float32)
        for k in range(3):
            acc = acc + p[:, k] * r[j, k]
        out[:, j] = (acc.astype(np.float16).astype(np.float32) + t[j]).astype(np.float16)
    return out


def add_error(pts_m: np.ndarray, pred: np.ndarray, gt: np.ndarray) -> float:
    """ADD, utils/metrics.py:205-218: float16 difference, float16 norm (squares rounded to float16, summed in float32,
    rounded, sqrt), mean with a float32 accumulator rounded to float16."""
    d = (_transform16(pts_m, pred) - _transform16(pts_m, gt)).astype(np.float32)
    sq = (d * d).astype(np.float16).astype(np.float32)
    s = (sq[:, 0] + (sq[:, 1] + sq[:, 2])).astype(np.float16).astype(np.float32)
    nrm = np.sqrt(s).astype(np.float16)
    return float(np.float16(np.float32(nrm.astype(np.float64).sum()) / np.float32(nrm.shape[0])))


def adds_error(pts_m: np.ndarray, pred: np.ndarray, gt: np.ndarray, block: int = 512) -> float:
    """ADD-S, utils/metrics.py:220-234: mean float64 distance from every float16 predicted point to its nearest float16
    ground-truth point (the reference asks a KDTree; brute force gives the same minimum)."""
    a = _transform16(pts_m, pred).astype(np.float64)
    b = _transform16(pts_m, gt).astype(np.float64)
    best = np.empty(a.shape[0])
    for i in range(0, a.shape[0], block):
        d2 = ((a[i:i + block, None, :] - b[None, :, :]) ** 2).sum(-1)
        best[i:i + block] = np.sqrt(d2.min(1))
    return float(best.mean())


def _pose16_mm(pose: np.ndarray):
    p16 = np.asarray(pose).astype(np.float16)
    return p16[:3, :3].astype(np.float64), (p16[:3, 3] * np.float16(1000)).astype(np.float16).astype(np.float64)


def mssd_mspd(pts_mm: np.ndarray, syms: np.ndarray, pred: np.ndarray, gt: np.ndarray, K: np.ndarray):
    """``my_mssd`` / ``my_mspd`` (bop_toolkit_lib/pose_error.py:370-426) as the evaluator calls them (:258-266).
    ``syms [S,3,4]``.  Only ``pts_mm[:3]`` enter (module docstring)."""
    P = np.asarray(pts_mm, dtype=np.float64)[:3]
    Rp, tp = _pose16_mm(pred)
    Rg, tg = _pose16_mm(gt)
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    est = P @ Rp.T + tp
    proj_est = est @ K.T
    proj_est = proj_est[:, :2] / proj_est[:, 2:3]
    best_s, best_p = np.inf, np.inf
    for s in np.asarray(syms, dtype=np.float64):
        Rs, ts = Rg @ s[:, :3], Rg @ s[:, 3] + tg
        q = P @ Rs.T + ts
        best_s = min(best_s, np.linalg.norm(est - q, axis=1).max())
        pq = q @ K.T
        pq = pq[:, :2] / pq[:, 2:3]
        best_p = min(best_p, np.linalg.norm(proj_est - pq, axis=1).max())
    return float(best_s), float(best_p)


def pose_errors(models: dict, syms: dict, cls_ids, pred: np.ndarray, gt: np.ndarray, cams: np.ndarray) -> np.ndarray:
    """Same contract as ``oryon_b200.utils.evaluator.cuda_pose_errors``: ``[P,6]`` float64 rows
    ``(R error deg, T error cm, ADD or ADD-S in m, used ADD-S flag, MSSD mm, MSPD px)``."""
    out = np.zeros((len(cls_ids), 6))
    for i, cid in enumerate(cls_ids):
        pts, sym = np.asarray(models[cid]["pts"], dtype=np.float64), np.asarray(syms[cid], dtype=np.float64)
        r, t = rt_errors(pred[i], gt[i])
        sym_obj = sym.shape[0] > 1
        a = adds_error(pts / 1000., pred[i], gt[i]) if sym_obj else add_error(pts / 1000., pred[i], gt[i])
        ms, mp = mssd_mspd(pts, sym, pred[i], gt[i], cams[i])
        out[i] = [r, t, a, float(sym_obj), ms, mp]
    return out
