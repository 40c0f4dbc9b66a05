"""CPU restatement of the VSD part of ``Evaluator.register_eval`` (reference utils/evaluator.py:279-286):
``bop_toolkit_lib/pose_error.py:17-96`` (``vsd``), ``visibility.py:9-75`` and ``misc.py:110-163`` (distance images), plus a
depth rasteriser standing in for ``bop_toolkit_lib/renderer_vispy.py``.  TEST INFRASTRUCTURE (checker for ``oryon_eval_vsd``).

The error arithmetic is pinned: ``oracle/make_golden_vsd.py`` runs the reference's own ``vsd()`` and ``Evaluator`` on depth
images produced by ``rasterize_depth`` below (handed over through a stand-in renderer object).  The RASTERISER is not
pinned -- the reference renders with OpenGL (vispy), which cannot run here: **parity of the rendering step is unpinned**.
It follows the reference's conventions read off its code: the projection of ``_calc_calib_proj`` (renderer_vispy.py:186-231)
puts the sample of output pixel (r, c) at image coordinates (c + 0.5, r + 0.5); no face culling, nearest surface wins (:549);
the value is the eye-space depth of the surface at the sample (the z-buffer read-back of :605-615 inverts the perspective
depth mapping), 0 where nothing is hit; float32.
"""
import numpy as np


def rasterize_depth(pts: np.ndarray, faces: np.ndarray, R: np.ndarray, t: np.ndarray, fx, fy, cx, cy, H: int, W: int) -> np.ndarray:
    """Depth image (float32 ``[H,W]``, same unit as ``pts`` / ``t``) of the mesh in pose ``(R, t)``.  float64 arithmetic in a
    fixed operation order (no fused multiply-add) so that the CUDA rasteriser makes bit-identical coverage decisions."""
    P = np.asarray(pts, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(3)
    cam = np.empty_like(P)
    for j in range(3):
        cam[:, j] = ((R[j, 0] * P[:, 0] + R[j, 1] * P[:, 1]) + R[j, 2] * P[:, 2]) + t[j]
    Z = cam[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        U = (fx * cam[:, 0]) / Z + cx
        V = (fy * cam[:, 1]) / Z + cy
    depth = np.full((H, W), np.inf, dtype=np.float32)
    for a, b, c in np.asarray(faces, dtype=np.int64):
        za, zb, zc = Z[a], Z[b], Z[c]
        if not (za > 0 and zb > 0 and zc > 0):
            continue
        ua, va, ub, vb, uc, vc = U[a], V[a], U[b], V[b], U[c], V[c]
        c0 = max(0, int(np.ceil(min(ua, ub, uc) - 0.5)))
        c1 = min(W - 1, int(np.floor(max(ua, ub, uc) - 0.5)))
        r0 = max(0, int(np.ceil(min(va, vb, vc) - 0.5)))
        r1 = min(H - 1, int(np.floor(max(va, vb, vc) - 0.5)))
        if c1 < c0 or r1 < r0:
            continue
        px = (np.arange(c0, c1 + 1, dtype=np.float64) + 0.5)[None, :]
        py = (np.arange(r0, r1 + 1, dtype=np.float64) + 0.5)[:, None]
        w0 = (ub - px) * (vc - py) - (uc - px) * (vb - py)
        w1 = (uc - px) * (va - py) - (ua - px) * (vc - py)
        w2 = (ua - px) * (vb - py) - (ub - px) * (va - py)
        inside = ((w0 >= 0) & (w1 >= 0) & (w2 >= 0)) | ((w0 <= 0) & (w1 <= 0) & (w2 <= 0))
        s = (w0 + w1) + w2
        inside &= s != 0
        if not inside.any():
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = ((w0 / za + w1 / zb) + w2 / zc) / s
            d = (1.0 / invz).astype(np.float32)
        win = depth[r0:r1 + 1, c0:c1 + 1]
        np.minimum(win, np.where(inside, d, np.float32(np.inf)), out=win)
    depth[np.isinf(depth)] = 0
    return depth


def dist_image(depth: np.ndarray, K: np.ndarray) -> np.ndarray:
    """misc.py:137-163."""
    H, W = depth.shape
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    pre_x = (xs - K[0, 2]) / np.float64(K[0, 0])
    pre_y = (ys - K[1, 2]) / np.float64(K[1, 1])
    return np.sqrt(np.multiply(pre_x, depth) ** 2 + np.multiply(pre_y, depth) ** 2 + depth.astype(np.float64) ** 2)


def vsd_errors(depth_est: np.ndarray, depth_gt: np.ndarray, depth_test: np.ndarray, K: np.ndarray, delta: float, taus, diameter: float):
    """pose_error.py:41-96 with ``normalized_by_diameter=True``, ``cost_type='step'``, visibility mode 'bop19'."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    d_test, d_gt, d_est = dist_image(depth_test, K), dist_image(depth_gt, K), dist_image(depth_est, K)

    def visib(d_model):
        diff = d_model.astype(np.float32) - d_test.astype(np.float32)
        return np.logical_and(np.logical_or(diff <= delta, d_test == 0), d_model > 0)

    v_gt = visib(d_gt)
    v_est = np.logical_or(visib(d_est), np.logical_and(v_gt, d_est > 0))
    inter, union = np.logical_and(v_gt, v_est), np.logical_or(v_gt, v_est)
    n_union = union.sum()
    n_comp = n_union - inter.sum()
    dists = np.abs(d_gt[inter] - d_est[inter]) / diameter
    if n_union == 0:
        return [1.0] * len(taus)
    return [(np.sum(dists >= tau) + n_comp) / float(n_union) for tau in taus]


class OracleRenderer:
    """Stand-in with the two methods of ``RendererVispy`` the evaluator uses (renderer_vispy.py:311, :512)."""

    def __init__(self, width, height, mode="depth"):
        self.width, self.height, self.models = width, height, {}

    def my_add_object(self, model: dict, obj_id):
        self.models[obj_id] = model

    def render_object(self, obj_id, R, t, fx, fy, cx, cy, clear=True):
        m = self.models[obj_id]
        R32, t32 = np.asarray(R).astype(np.float32), np.asarray(t).astype(np.float32).reshape(3)   # mat_view_cv is float32 (:520-521)
        return {"depth": rasterize_depth(m["pts"], m["faces"], R32, t32, fx, fy, cx, cy, self.height, self.width)}
