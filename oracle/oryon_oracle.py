"""CPU oracle for the Oryon post-network hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as
the checker / timed CPU baseline.  ``oryon_b200/`` never imports it: the product path fails loudly
when the CUDA library is missing.

What it is: a restatement, in plain CPU PyTorch (the reference itself is CPU PyTorch for this part
of the path, ``configs/config.yaml:7  corrs_device: cpu``), of the reference's algorithm, each
function citing the reference ``file:line`` it follows.  It is float32 where the reference is
float32 and int64 where the reference is int64.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference code itself*, produced in the build container
by ``oracle/make_golden.py`` (which imports the unmodified modules from /root/reference) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function below
against those fixtures bit-for-bit (indices) / to 0 ulp or stated tolerance (floats).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

# --------------------------------------------------------------------------------------------
# a7  mask post-processing  (reference losses.py:56-60, utils/metrics.py:18-40)
# --------------------------------------------------------------------------------------------


def predicted_mask(mask_logits: Tensor, mask_th: float = 0.5) -> Tensor:
    """``torch.where(sigmoid(logits) > th, 1, 0)`` on ``[B,1,H,W]`` logits -> int64 ``[B,H,W]``
    (reference losses.py:56-59)."""
    logits = mask_logits.squeeze(1)
    return torch.where(torch.sigmoid(logits) > mask_th, 1, 0)


def mask_iou(gt: Tensor, pred: Tensor) -> Tensor:
    """IoU per sample of two ``[B,H,W]`` masks (reference utils/metrics.py:18-40).
    0/0 yields NaN exactly as the reference's integer/integer true division does."""
    b = gt.shape[0]
    g, p = gt.reshape(b, -1), pred.reshape(b, -1)
    union = torch.logical_or(g, p).sum(1)
    inter = torch.logical_and(g, p).sum(1)
    return inter / union


def resize_mask_nearest(mask: Tensor, size: Tuple[int, int]) -> Tensor:
    """Oracle/ovseg mask path: nearest resize of a ``[H,W]`` mask to the feature-map size, cast to
    int32 (reference pipeline.py:380-386 and :409-412)."""
    m = mask.clone().to(torch.float)[None, None]
    return F.interpolate(m, tuple(size), mode="nearest").squeeze().to(torch.int)


# --------------------------------------------------------------------------------------------
# a8  matching  (reference utils/pcd.py:22-33, 177-216; utils/misc.py:242-254)
# --------------------------------------------------------------------------------------------


def inv_norm_cosine(A: Tensor, B: Tensor, row_chunk: int = 0) -> Tensor:
    """``0.5 * (1 - cos(A_i, B_j))`` for all pairs, ``[N1,D] x [N2,D] -> [N1,N2]`` float32
    (reference utils/pcd.py:28-29, the broadcast ``cosine_similarity(A[:,None], B[None], dim=2)``).

    ``row_chunk`` only bounds the size of the broadcast temporary (the reference materialises
    N1 x N2 x D); every output element is produced by the same call on the same operands, so the
    values are identical to the un-chunked form (checked in tests/test_oracle_golden.py)."""
    if row_chunk <= 0 or A.shape[0] <= row_chunk:
        return 0.5 * (-1 * F.cosine_similarity(A.unsqueeze(1), B.unsqueeze(0), dim=2) + 1)
    out = torch.empty(A.shape[0], B.shape[0], dtype=A.dtype)
    for r0 in range(0, A.shape[0], row_chunk):
        a = A[r0:r0 + row_chunk]
        out[r0:r0 + row_chunk] = 0.5 * (-1 * F.cosine_similarity(a.unsqueeze(1), B.unsqueeze(0), dim=2) + 1)
    return out


def sample_select(n_items: int, n: int, generator: Optional[torch.Generator] = None) -> Tensor:
    """Exactly ``n`` indices out of ``n_items``; replacement iff ``n > n_items``; float64 uniform
    weights on the CPU generator (reference utils/misc.py:242-254)."""
    w = torch.ones(n_items, dtype=torch.float64)
    return torch.multinomial(w, n, replacement=(n > n_items), generator=generator)


def match_rows(roi_feats1: Tensor, roi_feats2: Tensor, row_chunk: int = 256) -> Tuple[Tensor, Tensor]:
    """Row-wise nearest neighbour under ``inv_norm_cosine``: ``(min_dist f32[N1], argmin i64[N1])``
    (reference utils/pcd.py:202-204).  ``argmin`` keeps the first minimum (lowest query index)."""
    n1 = roi_feats1.shape[0]
    min_dist = torch.empty(n1, dtype=torch.float32)
    arg = torch.empty(n1, dtype=torch.int64)
    step = row_chunk if row_chunk > 0 else max(n1, 1)
    for r0 in range(0, n1, step):
        d = inv_norm_cosine(roi_feats1[r0:r0 + step], roi_feats2)
        min_dist[r0:r0 + step] = torch.amin(d, dim=1)
        arg[r0:r0 + step] = torch.argmin(d, dim=1)
    return min_dist, arg


def nn_correspondences(feats1: Tensor, feats2: Tensor, mask1: Tensor, mask2: Tensor,
                       threshold: float, max_corrs: int, subsample_source: Optional[int],
                       generator: Optional[torch.Generator] = None,
                       return_debug: bool = False):
    """Correspondences ``(y1,x1,y2,x2)`` between two ``[D,H,W]`` feature maps, CPU float32 branch
    (reference utils/pcd.py:177-216 with ``corrs_device='cpu'``).

    Order of random draws on the CPU generator -- (1) source subsample iff ``N1 > subsample_source``
    (:187-190), (2) the final ``max_corrs`` draw (:211) -- is the reference's; ``None`` is returned
    when at most one row passes the threshold (:206, :213-214)."""
    roi1 = torch.nonzero(mask1 == 1)
    roi2 = torch.nonzero(mask2 == 1)
    if subsample_source is not None and roi1.shape[0] > subsample_source:
        roi1 = roi1[sample_select(roi1.shape[0], subsample_source, generator)]
    f1 = feats1[:, roi1[:, 0], roi1[:, 1]].T.to(torch.float32)
    f2 = feats2[:, roi2[:, 0], roi2[:, 1]].T.to(torch.float32)
    if f1.shape[0] == 0 or f2.shape[0] == 0:
        # the reference would raise inside amin on an empty dim; callers gate on
        # is_detection_valid (pipeline.py:316) so this branch is never reached there.
        debug = dict(roi1=roi1, roi2=roi2, min_dist=torch.empty(0), nn_idx=torch.empty(0, dtype=torch.int64))
        return (None, debug) if return_debug else None
    min_dist, nn_idx = match_rows(f1, f2)
    valid = torch.nonzero(min_dist < threshold).squeeze(1)
    corrs = None
    if valid.shape[0] > 1:
        final = torch.cat((roi1[valid], roi2[nn_idx][valid]), dim=1)
        corrs = final[sample_select(final.shape[0], max_corrs, generator)]
    if return_debug:
        return corrs, dict(roi1=roi1, roi2=roi2, min_dist=min_dist, nn_idx=nn_idx, valid=valid)
    return corrs


# --------------------------------------------------------------------------------------------
# a9  coordinate scaling / bounds  (reference utils/coordinates.py:5-13, 36-47)
# --------------------------------------------------------------------------------------------


def nn_correspondences_kp(feats1: Tensor, feats2: Tensor, kp1: Tensor, kp2: Tensor, threshold: float, max_corrs: int,
                          max_source: Optional[int] = None, keep_empty: bool = False,
                          generator: Optional[torch.Generator] = None, return_debug: bool = False):
    """Matches between two descriptor sets ``[N,D]`` with key points ``[N,2]``, as the reference's key-point baselines
    compute them (scripts/evaluation/sift_nocs.py:25-45; sift_toyl.py:25-51 = ``max_source=1000, keep_empty=True``: the
    source subsample when N1 > 1000 (:32-35) is the first draw, and an empty match set skips the final draw (:45-47))."""
    if max_source is not None and feats1.shape[0] > max_source:
        idxs = sample_select(feats1.shape[0], max_source, generator)
        feats1, kp1 = feats1[idxs], kp1[idxs]
    min_dist, nn_idx = match_rows(feats1, feats2)
    valid = torch.nonzero(min_dist < threshold).squeeze(1)
    final_corrs = torch.cat((kp1[valid], kp2[nn_idx][valid]), dim=1)
    if not (keep_empty and final_corrs.shape[0] == 0):
        final_corrs = final_corrs[sample_select(final_corrs.shape[0], max_corrs, generator)]
    if return_debug:
        return final_corrs, dict(nn_idx=nn_idx, min_dist=min_dist, valid=valid, feats1=feats1, kp1=kp1)
    return final_corrs


def scale_coords(coords: Tensor, source_scale, target_scale) -> Tensor:
    """(y,x) * (target/source) in float32; the ratio is a Python float (reference
    utils/coordinates.py:5-13)."""
    out = coords.clone().to(torch.float32)
    out[:, 0] = out[:, 0] * (target_scale[0] / source_scale[0])
    out[:, 1] = out[:, 1] * (target_scale[1] / source_scale[1])
    return out


def get_valid_coords(coords: Tensor, bounds) -> Tensor:
    """``0 <= y < H and 0 <= x < W`` (reference utils/coordinates.py:36-47)."""
    ys, xs = coords[:, 0], coords[:, 1]
    return (xs >= 0) & (xs < bounds[1]) & (ys >= 0) & (ys < bounds[0])


# --------------------------------------------------------------------------------------------
# a10  2D -> 3D lifting  (reference utils/pcd.py:35-81, xy_idxs branch; pipeline.py:438-460)
# --------------------------------------------------------------------------------------------


def lift_pcd(depth: Tensor, camera: Tensor, xy_idxs: Tuple[Tensor, Tensor]) -> Tensor:
    """Pin-hole unprojection of selected pixels.  ``depth [H,W,1]``, ``camera [9]`` row-major K,
    ``xy_idxs = (x i64[n], y i64[n])`` -> ``[n,3]`` in depth units (reference utils/pcd.py:35-81).

    Type promotion follows the reference: x,y are float32, camera entries are 0-dim tensors of the
    camera dtype (float64 in the batch schema), so the products stay float32 for float depth and
    become float32 for integer depth too (0-dim tensors do not promote dimensioned ones)."""
    d = depth[:, :, 0]
    xmap, ymap = xy_idxs[0], xy_idxs[1]
    z = d[ymap, xmap]
    xf, yf = xmap.to(torch.float32), ymap.to(torch.float32)
    fx, fy, cx, cy = camera[0], camera[4], camera[2], camera[5]
    px = (xf - cx) * z / fx
    py = (yf - cy) * z / fy
    return torch.stack((px, py, z), dim=1)


def corrs_to_pcds(corrs: Tensor, depth_a: Tensor, depth_q: Tensor, camera_a: Tensor, camera_q: Tensor,
                  featmap_size: Tuple[int, int], size_a: Tuple[int, int], size_q: Tuple[int, int]
                  ) -> Tuple[Tensor, Tensor]:
    """From featmap-space correspondences ``[n,4]`` to two metric point sets ``[m,3]`` (metres):
    scale to raw-frame pixels, drop out-of-bounds rows, truncate to int64, lift, divide by 1000
    (reference pipeline.py:447-460)."""
    ca = scale_coords(corrs[:, :2].clone(), featmap_size, size_a)
    cq = scale_coords(corrs[:, 2:].clone(), featmap_size, size_q)
    ok = get_valid_coords(ca, size_a) & get_valid_coords(cq, size_q)
    ca, cq = ca[ok].to(torch.long), cq[ok].to(torch.long)
    pa = lift_pcd(depth_a.unsqueeze(-1), camera_a.reshape(9), (ca[:, 1], ca[:, 0])) / 1000.0
    pq = lift_pcd(depth_q.unsqueeze(-1), camera_q.reshape(9), (cq[:, 1], cq[:, 0])) / 1000.0
    return pa, pq


# --------------------------------------------------------------------------------------------
# a11  PointDSC registration, test mode, bs == 1
#      (reference models/pointdsc/PointDSC.py, models/pointdsc/common.py, utils/pointdsc/SE3.py,
#       utils/pointdsc/init.py)
# --------------------------------------------------------------------------------------------


def _conv1x1(w: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    """``nn.Conv1d(kernel_size=1)`` on ``[C_in, N]`` (W[C_out,C_in,1] @ x + b), evaluated through the
    same ATen kernel the reference's ``nn.Conv1d`` module dispatches to, so results are bit-identical."""
    return F.conv1d(x[None], w[name + ".weight"], w[name + ".bias"])[0]


def _bn_eval(w: Dict[str, Tensor], name: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """``nn.BatchNorm1d`` in eval mode on ``[C, N]``: ``(x - mean) / sqrt(var + eps) * g + b`` with the
    running statistics."""
    return F.batch_norm(x[None], w[name + ".running_mean"], w[name + ".running_var"],
                        w[name + ".weight"], w[name + ".bias"], False, 0.0, eps)[0]


def spatial_consistency(src: Tensor, tgt: Tensor, sigma_d: float) -> Tuple[Tensor, Tensor]:
    """SC matrix ``clamp(1 - (|si-sj| - |ti-tj|)^2 / sigma_d^2, 0)`` and the source distance matrix
    (reference PointDSC.py:150-153)."""
    sd = torch.norm(src[:, None, :] - src[None, :, :], dim=-1)
    td = torch.norm(tgt[:, None, :] - tgt[None, :, :], dim=-1)
    sc = torch.clamp(1.0 - (sd - td) ** 2 / sigma_d ** 2, min=0)
    return sc, sd


def nonlocal_net(w: Dict[str, Tensor], corr_pos: Tensor, sc: Tensor, num_layers: int, num_channels: int
                 ) -> Tensor:
    """NonLocalNet encoder: layer0, then ``num_layers`` x (PointCN conv+BN+ReLU, NonLocalBlock)
    (reference PointDSC.py:48-77 and :9-45).  ``corr_pos [N,6]`` -> features ``[C,N]``."""
    feat = _conv1x1(w, "encoder.layer0", corr_pos.T)
    for i in range(num_layers):
        p = f"encoder.blocks.PointCN_layer_{i}"
        feat = torch.relu(_bn_eval(w, p + ".1", _conv1x1(w, p + ".0", feat)))
        p = f"encoder.blocks.NonLocal_layer_{i}"
        q = _conv1x1(w, p + ".projection_q", feat)
        k = _conv1x1(w, p + ".projection_k", feat)
        v = _conv1x1(w, p + ".projection_v", feat)
        att = torch.einsum("co,ci->oi", q, k) / num_channels ** 0.5   # [N(o), N(i)], one head
        weight = torch.softmax(sc * att, dim=-1)
        msg = torch.einsum("oi,ci->co", weight, v)                 # [C, N(o)]
        msg = torch.relu(_bn_eval(w, p + ".fc_message.1", _conv1x1(w, p + ".fc_message.0", msg)))
        msg = torch.relu(_bn_eval(w, p + ".fc_message.4", _conv1x1(w, p + ".fc_message.3", msg)))
        msg = _conv1x1(w, p + ".fc_message.6", msg)
        feat = feat + msg
    return feat


def confidence_mlp(w: Dict[str, Tensor], feat: Tensor) -> Tensor:
    """Classification head 128->32->32->1 with ReLUs (reference PointDSC.py:107-113, :171)."""
    x = torch.relu(_conv1x1(w, "classification.0", feat))
    x = torch.relu(_conv1x1(w, "classification.2", x))
    return _conv1x1(w, "classification.4", x)[0]


def pick_seeds(src_dist: Tensor, scores: Tensor, radius: float, max_num: int) -> Tensor:
    """Parallel NMS seed selection (reference PointDSC.py:199-217): a point is a local maximum iff
    every point within ``radius`` has a score <= its own; seeds are the top ``max_num`` of
    ``score * is_local_max`` under ``argsort(descending=True)``."""
    rel = (scores[:, None] >= scores[None, :]) | (src_dist >= radius)
    is_max = rel.min(-1)[0].float()
    return torch.argsort(scores * is_max, descending=True)[:max_num]


def knn_feature(x: Tensor, k: int) -> Tensor:
    """k nearest neighbours in (normalised) feature space, self excluded by dropping the first of
    ``topk(k+1, largest=False)`` over ``2 - 2 x x^T`` (reference common.py:48-69)."""
    dist = 2 - 2 * (x @ x.T)
    return dist.topk(k=k + 1, dim=-1, largest=False)[1][:, 1:]


def leading_eigenvector(M: Tensor, num_iterations: int) -> Tensor:
    """Power iteration with the *global* ``allclose`` early exit over all seeds
    (reference PointDSC.py:338-358)."""
    v = torch.ones_like(M[:, :, 0:1])
    last = v
    for _ in range(num_iterations):
        v = torch.bmm(M, v)
        v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
        if torch.allclose(v, last):
            break
        last = v
    return v.squeeze(-1)


def rigid_transform_3d(A: Tensor, B: Tensor, weights: Optional[Tensor] = None) -> Tensor:
    """Weighted Kabsch, batched ``[b,n,3]`` -> ``[b,4,4]`` (reference common.py:7-45): centroids
    with +1e-6 in the denominator, ``H = Am^T diag(w) Bm``, ``U,S,V = svd(H)``,
    ``R = V diag(1,1,det(V U^T)) U^T``, ``t = cB - R cA``."""
    b = A.shape[0]
    if weights is None:
        weights = torch.ones_like(A[:, :, 0])
    weights = weights.clone()
    weights[weights < 0] = 0
    wsum = torch.sum(weights, dim=1, keepdim=True)[:, :, None] + 1e-6
    cA = torch.sum(A * weights[:, :, None], dim=1, keepdim=True) / wsum
    cB = torch.sum(B * weights[:, :, None], dim=1, keepdim=True) / wsum
    Am, Bm = A - cA, B - cB
    H = Am.permute(0, 2, 1) @ torch.diag_embed(weights) @ Bm
    U, _, V = torch.svd(H)
    d = torch.det(V @ U.permute(0, 2, 1))
    E = torch.eye(3)[None].repeat(b, 1, 1)
    E[:, -1, -1] = d
    R = V @ E @ U.permute(0, 2, 1)
    t = cB.permute(0, 2, 1) - R @ cA.permute(0, 2, 1)
    T = torch.eye(4)[None].repeat(b, 1, 1)
    T[:, :3, :3] = R
    T[:, :3, 3:4] = t
    return T


def seed_hypotheses(seeds: Tensor, feats_n: Tensor, src: Tensor, tgt: Tensor, *, k: int, sigma: float,
                    sigma_d: float, num_iterations: int, inlier_threshold: float
                    ) -> Tuple[Tensor, Tensor, Tensor]:
    """Per-seed local spectral matching + weighted Kabsch + hypothesis scoring
    (reference PointDSC.py:234-336, ``seed_as_center=False`` branch).
    Returns ``(seed_trans [S,4,4], fitness [S], best_trans [4,4])``."""
    n = feats_n.shape[0]
    k = min(k, n - 1)
    idx = knn_feature(feats_n, k)[seeds]                                   # [S,k]
    kf = feats_n[idx]                                                      # [S,k,C]
    fM = torch.clamp(1 - (1 - kf @ kf.permute(0, 2, 1)) / sigma ** 2, min=0)
    s_k, t_k = src[idx], tgt[idx]                                          # [S,k,3]
    sd = ((s_k[:, :, None, :] - s_k[:, None, :, :]) ** 2).sum(-1) ** 0.5
    td = ((t_k[:, :, None, :] - t_k[:, None, :, :]) ** 2).sum(-1) ** 0.5
    sM = torch.clamp(1 - (sd - td) ** 2 / sigma_d ** 2, min=0)
    M = fM * sM
    ar = torch.arange(k)
    M[:, ar, ar] = 0
    wgt = leading_eigenvector(M, num_iterations)
    wgt = wgt / (torch.sum(wgt, dim=-1, keepdim=True) + 1e-6)
    seed_trans = rigid_transform_3d(s_k, t_k, wgt)
    pred = torch.einsum("snm,mk->snk", seed_trans[:, :3, :3], src.T) + seed_trans[:, :3, 3:4]  # [S,3,N]
    l2 = torch.norm(pred.permute(0, 2, 1) - tgt[None], dim=-1)
    fitness = torch.mean((l2 < inlier_threshold).float(), dim=-1)
    best = fitness.argmax()
    return seed_trans, fitness, seed_trans[best]


def post_refinement(T: Tensor, src: Tensor, tgt: Tensor, inlier_threshold: float, max_iters: int = 20
                    ) -> Tensor:
    """Iterative re-weighted Kabsch on the current inlier set; stops when the inlier count repeats
    (reference PointDSC.py:403-438; thresholds list is ``[0.10]*20`` when
    ``inlier_threshold == 0.10`` else ``[1.2]*20``, :415-418)."""
    th = 0.10 if inlier_threshold == 0.10 else 1.2
    prev = 0
    for _ in range(max_iters):
        warped = (T[:3, :3] @ src.T + T[:3, 3:4]).T
        l2 = torch.norm(warped - tgt, dim=-1)
        inl = l2 < th
        cnt = int(inl.sum())
        if abs(cnt - prev) < 1:
            break
        prev = cnt
        T = rigid_transform_3d(src[None, inl], tgt[None, inl], (1 / (1 + (l2 / th) ** 2))[None, inl])[0]
    return T


def pointdsc_pose(w: Dict[str, Tensor], cfg: Dict, pcd1: Tensor, pcd2: Tensor, return_debug: bool = False):
    """``get_pointdsc_pose`` + ``PointDSC.forward`` in test mode (reference utils/pointdsc/init.py:10-29,
    PointDSC.py:128-197).  ``w`` is the model ``state_dict``; ``cfg`` carries ``num_layers,
    num_channels, num_iterations, ratio, sigma_d, k, inlier_threshold (-> nms_radius, init.py:49)``;
    the model's own ``inlier_threshold`` stays at the constructor default 0.10 (PointDSC.py:87)."""
    pcd1, pcd2 = pcd1.float(), pcd2.float()
    corr_pos = torch.cat([pcd1, pcd2], dim=-1)
    corr_pos = corr_pos - corr_pos.mean(0)
    sigma_d = float(w["sigma_spat"][0]) if "sigma_spat" in w else float(cfg["sigma_d"])
    sigma = float(w["sigma"][0]) if "sigma" in w else 1.0
    n = pcd1.shape[0]
    sc, sd = spatial_consistency(pcd1, pcd2, sigma_d)
    feat = nonlocal_net(w, corr_pos, sc, cfg["num_layers"], cfg["num_channels"])      # [C,N]
    feats_n = F.normalize(feat.T, p=2, dim=-1)
    conf = confidence_mlp(w, feat)
    seeds = pick_seeds(sd, conf, cfg["inlier_threshold"], int(n * cfg["ratio"]))
    seed_trans, fitness, best = seed_hypotheses(
        seeds, feats_n, pcd1, pcd2, k=cfg["k"], sigma=sigma, sigma_d=sigma_d,
        num_iterations=cfg["num_iterations"], inlier_threshold=0.10)
    final = post_refinement(best, pcd1, pcd2, 0.10)
    if return_debug:
        return final, dict(sc=sc, feat=feat, conf=conf, seeds=seeds, seed_trans=seed_trans,
                           fitness=fitness, initial=best)
    return final


# --------------------------------------------------------------------------------------------
# a12  the per-pair loop of test_step  (reference pipeline.py:313-355) on given network outputs
# --------------------------------------------------------------------------------------------


def post_network_step(outputs: Dict[str, Tensor], batch: Dict, pointdsc_w: Dict[str, Tensor], pointdsc_cfg: Dict, *,
                      mask_mode: str = "predicted", mask_th: float = 0.5, dist_th: float = 0.25, n_corrs: int = 500,
                      src_sampling: Optional[int] = 5000, featmap_size: Tuple[int, int] = (192, 192)):
    """Everything ``test_step`` does after ``model.forward`` (reference pipeline.py:311-355), pair by pair and in
    the reference's order (so the CPU generator is consumed exactly as the reference consumes it): mask
    post-processing and IoU (losses.py:52-60), validity (:372-395), ``nn_correspondences`` (:397-427),
    scale / lift (:438-460), ``get_pointdsc_pose`` (:468), ``pred_q = pred_pose @ anchor_pose`` (:320), identity
    pose on failure (:335-350).  ``outputs`` holds CPU tensors."""
    B = outputs["featmap_a"].shape[0]
    res = {}
    for v, key in (("a", "anchor"), ("q", "query")):
        gt = batch[key]["mask"]
        gt_c = F.interpolate(gt.unsqueeze(1).float(), tuple(featmap_size), mode="nearest").squeeze(1)
        res["mask_" + v] = predicted_mask(outputs["mask_" + v], mask_th)
        res["iou_" + v] = mask_iou(gt_c, res["mask_" + v])
    rows = []
    for b in range(B):
        if mask_mode != "predicted":
            ma = resize_mask_nearest(batch["anchor"]["mask"][b], featmap_size)
            mq = resize_mask_nearest(batch["query"]["mask"][b], featmap_size)
        else:
            ma, mq = res["mask_a"][b], res["mask_q"][b]
        valid = int(torch.count_nonzero(ma == 1)) > 0 and int(torch.count_nonzero(mq == 1)) > 0
        pose, status, corrs = torch.eye(4), "invalid_mask", None
        if valid:
            corrs = nn_correspondences(outputs["featmap_a"][b].clone(), outputs["featmap_q"][b].clone(), ma, mq, dist_th, n_corrs,
                                       src_sampling)
            status = "no_corrs"
            if corrs is not None:
                HA, WA = (int(x) for x in batch["anchor"]["sizes"][b])
                HQ, WQ = (int(x) for x in batch["query"]["sizes"][b])
                pa, pq = corrs_to_pcds(corrs, batch["anchor"]["orig_depth"][b].squeeze(), batch["query"]["orig_depth"][b].squeeze(),
                                       batch["anchor"]["camera"][b].reshape(9), batch["query"]["camera"][b].reshape(9), featmap_size,
                                       (HA, WA), (HQ, WQ))
                pose, status = pointdsc_pose(pointdsc_w, pointdsc_cfg, pa, pq).to(torch.float32), "ok"
        pred_q = pose @ batch["anchor"]["pose"][b].to(torch.float32) if status == "ok" else None
        rows.append(dict(pred_pose_rel=pose, pred_pose=pred_q, iou_a=float(res["iou_a"][b]), iou_q=float(res["iou_q"][b]),
                         status=status, corrs=corrs))
    return rows
