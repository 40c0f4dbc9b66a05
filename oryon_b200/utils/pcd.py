"""Mirror of the reference ``utils/pcd.py`` for the hot path: ``nn_correspondences`` (:177-216) and
``lift_pcd`` (:35-81), same names, argument meaning and error behaviour, computed by
liboryon_b200.so on the GPU.

What stays in torch is what the reference itself leaves to the *caller's* generator and ordering:
the two ``torch.multinomial`` draws (utils/misc.py:242-254) and row selection by the drawn indices.
ROI enumeration (``torch.nonzero(mask == 1)`` order), the all-pairs inverted-cosine distance, the
row argmin and the threshold are computed by the library.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from .. import _lib
from .._torch_glue import as_device, device_of, ptr, stream_ptr

_DEPTH_DTYPES = {torch.int32: _lib.DEPTH_I32, torch.float32: _lib.DEPTH_F32, torch.int16: _lib.DEPTH_I16,
                 torch.uint16: _lib.DEPTH_U16}


def torch_sample_select(t: Tensor, n: int) -> Tensor:
    """Exactly ``n`` row indices of ``t``; replacement only if ``n > N`` (reference utils/misc.py:242-254).
    The draw uses the default generator of ``t.device`` exactly as the reference's call does."""
    N = t.shape[0]
    uniform_dist = torch.ones(N, dtype=float).to(t.device)
    return torch.multinomial(uniform_dist, n, replacement=bool(n > N)).to(t.device)


# ------------------------------------------------------------------------------------------------
# library calls
# ------------------------------------------------------------------------------------------------


def mask_to_roi(mask: Tensor, value: int = 1) -> Tuple[Tensor, Tensor]:
    """``[B,H,W]`` (or ``[H,W]``) integer masks on the GPU -> ``(pixel_ids int32 [B,HW], counts int32 [B])``;
    ``pixel_ids[b,:counts[b]]`` are the flattened ``y*W+x`` of ``mask[b] == value`` in ``torch.nonzero`` order."""
    dev = device_of(mask)
    m = as_device(mask, dev, torch.int32)
    if m.dim() == 2:
        m = m[None]
    B, HW = m.shape[0], m.shape[1] * m.shape[2]
    roi = torch.empty(B, HW, dtype=torch.int32, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.oryon_mask_to_roi(_lib.handle(dev.index), ptr(m), B, HW, int(value), ptr(roi), ptr(cnt), stream_ptr(dev)))
    return roi, cnt


def match_nn(feat_a: Tensor, feat_q: Tensor, roi_a: Optional[Tensor] = None, roi_q: Optional[Tensor] = None,
             n_a: Optional[Sequence[int]] = None, n_q: Optional[Sequence[int]] = None, *,
             mode: int = _lib.MATCH_TC_REFINED, out: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Row-wise nearest neighbour under ``0.5*(1-cos)`` for B pairs at once (``oryon_match_nn``).

    ``feat_a/q``: float32 ``[B,D,H,W]`` (or ``[B,D,HW]``) on the GPU.  ``roi_x``: int32 ``[B,cap]`` pixel-id
    lists with host lengths ``n_x`` or ``None`` for every pixel.  Returns ``(idx int32 [B,cap_a],
    dist float32 [B,cap_a])``: position in the pair's query list of the nearest neighbour (lowest on
    ties) and its distance; ``-1`` / ``inf`` beyond ``n_a[b]``."""
    dev = device_of(feat_a, feat_q)
    fa, fq = as_device(feat_a, dev, torch.float32), as_device(feat_q, dev, torch.float32)
    if fa.dim() < 3 or fq.dim() < 3 or fa.shape[:2] != fq.shape[:2]:
        raise ValueError(f"match_nn: feature maps must be [B,D,...] with equal B and D, got {tuple(fa.shape)} / {tuple(fq.shape)}")
    B, D = fa.shape[0], fa.shape[1]
    hw_a, hw_q = fa[0, 0].numel(), fq[0, 0].numel()
    if (roi_a is None) != (n_a is None) or (roi_q is None) != (n_q is None):
        raise ValueError("match_nn: roi_x and n_x go together")
    cap_a = hw_a if roi_a is None else roi_a.shape[1]
    cap_q = hw_q if roi_q is None else roi_q.shape[1]
    if roi_a is not None:
        roi_a = as_device(roi_a, dev, torch.int32)
    if roi_q is not None:
        roi_q = as_device(roi_q, dev, torch.int32)
    na = None if n_a is None else (ctypes.c_int32 * B)(*[int(v) for v in n_a])
    nq = None if n_q is None else (ctypes.c_int32 * B)(*[int(v) for v in n_q])
    if out is None:
        idx = torch.empty(B, cap_a, dtype=torch.int32, device=dev)
        dist = torch.empty(B, cap_a, dtype=torch.float32, device=dev)
    else:
        idx, dist = out
        if idx.shape != (B, cap_a) or dist.shape != (B, cap_a) or idx.dtype != torch.int32 or dist.dtype != torch.float32 \
                or idx.device != dev or dist.device != dev or not idx.is_contiguous() or not dist.is_contiguous():
            raise ValueError("match_nn: bad `out` buffers")
    lib = _lib.load()
    _lib.check(lib.oryon_match_nn(_lib.handle(dev.index), ptr(fa), ptr(fq), B, D, hw_a, hw_q, ptr(roi_a), ptr(roi_q), na, nq,
                                  cap_a, cap_q, int(mode), ptr(idx), ptr(dist), stream_ptr(dev)))
    return idx, dist


def match_nn_streamed(feat_a: Tensor, feat_q: Tensor, *, chunk_pairs: int = 4, out: Optional[Tuple[Tensor, Tensor]] = None
                      ) -> Tuple[Tensor, Tensor]:
    """Dense ``match_nn`` for HOST feature maps ``[B,D,H,W]`` (pinned memory for full speed): the batch is cut into
    chunks of ``chunk_pairs`` pairs whose host-to-device copies run on a side stream, double buffered, while the
    previous chunk is being matched -- the PCIe transfer (the slower of the two) hides the kernels.  Returns
    device ``(idx int32 [B,HW], dist float32 [B,HW])``; the work is ordered on the current stream."""
    dev = device_of()
    if feat_a.is_cuda or feat_q.is_cuda:
        return match_nn(feat_a, feat_q, out=out)
    B, D = feat_a.shape[:2]
    hw = feat_a[0, 0].numel()
    if out is None:
        out = (torch.empty(B, hw, dtype=torch.int32, device=dev), torch.empty(B, hw, dtype=torch.float32, device=dev))
    idx, dist = out
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)
    bufs = [(torch.empty(chunk_pairs, *feat_a.shape[1:], dtype=torch.float32, device=dev),
             torch.empty(chunk_pairs, *feat_q.shape[1:], dtype=torch.float32, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    starts = list(range(0, B, chunk_pairs))
    side.wait_stream(cur)

    def upload(i):
        b0, k = starts[i], i % 2
        n = min(chunk_pairs, B - b0)
        with torch.cuda.stream(side):
            if i >= 2:
                side.wait_event(freed[k])
            bufs[k][0][:n].copy_(feat_a[b0:b0 + n], non_blocking=True)
            bufs[k][1][:n].copy_(feat_q[b0:b0 + n], non_blocking=True)
            ready[k].record(side)

    upload(0)
    for i, b0 in enumerate(starts):
        if i + 1 < len(starts):
            upload(i + 1)
        k, n = i % 2, min(chunk_pairs, B - b0)
        cur.wait_event(ready[k])
        match_nn(bufs[k][0][:n], bufs[k][1][:n], out=(idx[b0:b0 + n], dist[b0:b0 + n]))
        freed[k].record(cur)
    for a, q in bufs:
        a.record_stream(cur), q.record_stream(cur)
    return idx, dist


def match_last_stats(device: Optional[torch.device] = None) -> dict:
    dev = device or device_of()
    s = (ctypes.c_int64 * 4)()
    _lib.check(_lib.load().oryon_match_last_stats(_lib.handle(dev.index), s, stream_ptr(dev)))
    return dict(rows_refined=s[0], chunks_rescored=s[1], rows_overflowed=s[2], kernels_launched=s[3])


HIST_CHUNK_BINS = [str(i) for i in range(17)] + ["overflow"]
HIST_COLUMN_BINS = ["1", "2", "3-4", "5-8", "9-16", "17-32", "33-64", "65+"]


def match_set_hist(enable: bool, device: Optional[torch.device] = None) -> None:
    """Diagnostic switch (``oryon_match_set_hist``): record candidate-list length histograms in the re-scoring pass."""
    dev = device or device_of()
    _lib.check(_lib.load().oryon_match_set_hist(_lib.handle(dev.index), int(bool(enable))))


def match_list_hist(device: Optional[torch.device] = None) -> dict:
    """Histograms of the last ``match_nn`` call (``oryon_match_list_hist``): anchor rows by candidate chunks kept by the
    tensor-core pass and by candidate columns re-scored in float32."""
    dev = device or device_of()
    h = (ctypes.c_int64 * 26)()
    _lib.check(_lib.load().oryon_match_list_hist(_lib.handle(dev.index), h, stream_ptr(dev)))
    return dict(rows_by_chunks=dict(zip(HIST_CHUNK_BINS, [int(v) for v in h[:18]])),
                rows_by_columns=dict(zip(HIST_COLUMN_BINS, [int(v) for v in h[18:26]])))


# ------------------------------------------------------------------------------------------------
# reference interface
# ------------------------------------------------------------------------------------------------


def nn_correspondences(feats1: Tensor, feats2: Tensor, mask1: Tensor, mask2: Tensor, threshold: float, max_corrs: int,
                       subsample_source: Optional[int], corrs_device: str = "cpu", *, return_debug: bool = False):
    """Finds matches between two ``[D,H,W]`` feature maps; returns ``int64 [max_corrs,4]`` rows
    ``(y1,x1,y2,x2)`` on ``feats1.device`` or ``None`` when at most one row passes ``threshold``
    (reference utils/pcd.py:177-216).

    ``corrs_device`` keeps its one observable meaning, the device whose default generator serves the two
    ``multinomial`` draws ('cpu' in the reference configuration, configs/config.yaml:7); the distance
    arithmetic always runs on the GPU and equals a float32 evaluation (ORYON_MATCH_TC_REFINED), which is
    what the reference computes for 'cpu' and is more precise than its float16 'cuda' branch."""
    orig_device = feats1.device
    dev = device_of(feats1, feats2, mask1, mask2)
    if feats1.dim() != 3 or feats2.dim() != 3 or feats1.shape[0] != feats2.shape[0]:
        raise ValueError("nn_correspondences: feature maps must be [D,H,W] with equal D")
    H1, W1 = feats1.shape[1:]
    H2, W2 = feats2.shape[1:]
    pix1, c1 = mask_to_roi(as_device(mask1, dev).reshape(1, H1, W1))
    pix2, c2 = mask_to_roi(as_device(mask2, dev).reshape(1, H2, W2))
    n1, n2 = (int(v) for v in torch.stack((c1[0], c2[0])).tolist())  # the reference syncs here too (nonzero)
    pix1, pix2 = pix1[0, :n1], pix2[0, :n2]

    if subsample_source is not None and n1 > subsample_source:
        probe = torch.empty(n1, 0, device=corrs_device)
        idxs = torch_sample_select(probe, subsample_source)
        pix1 = pix1[idxs.to(dev)]
        n1 = int(subsample_source)
    if n1 == 0 or n2 == 0:
        # torch.amin over an empty dimension raises in the reference (callers gate on is_detection_valid)
        raise RuntimeError("nn_correspondences: empty mask (amin over an empty dimension)")

    idx, dist = match_nn(feats1[None], feats2[None], pix1[None].contiguous(), pix2[None].contiguous(), [n1], [n2])
    idx, dist = idx[0, :n1], dist[0, :n1]
    valid = torch.nonzero(dist < threshold).squeeze(1)
    final_corrs = None
    if valid.shape[0] > 1:
        p1 = pix1[valid].long()
        p2 = pix2[idx[valid].long()].long()
        final = torch.stack((p1 // W1, p1 % W1, p2 // W2, p2 % W2), dim=1)
        probe = torch.empty(final.shape[0], 0, device=corrs_device)
        sel = torch_sample_select(probe, max_corrs)
        final_corrs = final[sel.to(dev)].to(orig_device)
    if return_debug:
        return final_corrs, dict(pix1=pix1, pix2=pix2, nn_idx=idx, min_dist=dist, valid=valid)
    return final_corrs


def nn_correspondences_kp(feats1: Tensor, feats2: Tensor, kp1: Tensor, kp2: Tensor, threshold: float, max_corrs: int,
                          max_source: Optional[int] = None, keep_empty: bool = False, *, return_debug: bool = False):
    """Matches between two descriptor SETS ``[N1,D]`` / ``[N2,D]`` with key-point coordinates ``kp1 [N1,2]`` / ``kp2 [N2,2]``:
    rows ``(kp1, kp2[nearest])`` of the pairs whose ``0.5*(1-cos)`` is below ``threshold``, resampled to exactly
    ``max_corrs`` rows -- the function of the same purpose in the reference's key-point baselines
    (scripts/evaluation/sift_nocs.py:25-45; sift_toyl.py:25-51 with ``max_source=1000, keep_empty=True``: more than
    ``max_source`` source descriptors are subsampled first, and an empty match set is returned as ``[0,4]`` instead of
    reaching ``torch.multinomial``, which raises on it).  The draws use the default generator of ``feats1.device`` as the
    reference's ``torch_sample_select`` calls do; the distances and the row argmin run in ``oryon_match_nn``."""
    if feats1.dim() != 2 or feats2.dim() != 2 or feats1.shape[1] != feats2.shape[1]:
        raise ValueError("nn_correspondences_kp: descriptor sets must be [N,D] with equal D")
    if kp1.shape[0] != feats1.shape[0] or kp2.shape[0] != feats2.shape[0]:
        raise ValueError("nn_correspondences_kp: one key point per descriptor")
    orig_device = feats1.device
    if max_source is not None and feats1.shape[0] > max_source:
        idxs = torch_sample_select(feats1, max_source)
        feats1, kp1 = feats1[idxs], kp1[idxs]
    n1, n2 = int(feats1.shape[0]), int(feats2.shape[0])
    if n1 == 0 or n2 == 0:
        raise RuntimeError("nn_correspondences_kp: empty descriptor set (amin over an empty dimension)")
    dev = device_of(feats1, feats2)
    # both sets as [1,D,cap] maps of one width, addressed through explicit position lists
    cap = max(n1, n2)
    fa = torch.zeros(1, feats1.shape[1], cap, dtype=torch.float32, device=dev)
    fq = torch.zeros(1, feats1.shape[1], cap, dtype=torch.float32, device=dev)
    fa[0, :, :n1] = as_device(feats1, dev, torch.float32).T
    fq[0, :, :n2] = as_device(feats2, dev, torch.float32).T
    pos = torch.arange(cap, dtype=torch.int32, device=dev)[None].contiguous()
    idx, dist = match_nn(fa, fq, pos, pos, [n1], [n2])
    idx, dist = idx[0, :n1], dist[0, :n1]
    valid = torch.nonzero(dist < threshold).squeeze(1)
    nearest = idx.long()[valid]
    final_corrs = torch.cat((kp1[valid.to(kp1.device)], kp2[nearest.to(kp2.device)].to(kp1.device)), dim=1).to(orig_device)
    if not (keep_empty and final_corrs.shape[0] == 0):
        final_corrs = final_corrs[torch_sample_select(final_corrs, max_corrs)]
    if return_debug:
        return final_corrs, dict(nn_idx=idx, min_dist=dist, valid=valid)
    return final_corrs


def lift_pcd(depth: Tensor, camera: Tensor, xy_idxs: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
    """Pin-hole lifting of a depth image ``[H,W,C]`` to a point cloud ``[n,3]`` in depth units (reference
    utils/pcd.py:35-81).  With ``xy_idxs = (x, y)`` only those pixels are lifted; without, every pixel in
    row-major order.  Extra channels (RGB point clouds, C > 1) are not part of the hot path."""
    if depth.dim() != 3:
        raise ValueError("lift_pcd: depth must be [H,W,C]")
    H, W, C = depth.shape
    if C != 1:
        raise NotImplementedError("lift_pcd: only single-channel depth is on the inference path")
    dev = device_of(depth)
    d = depth[:, :, 0]
    if d.dtype not in _DEPTH_DTYPES:
        d = d.to(torch.float32)
    d = as_device(d, dev)
    if xy_idxs is None:
        xs = torch.arange(W, device=dev, dtype=torch.int64).repeat(H)
        ys = torch.arange(H, device=dev, dtype=torch.int64).repeat_interleave(W)
    else:
        xs, ys = as_device(xy_idxs[0], dev, torch.int64), as_device(xy_idxs[1], dev, torch.int64)
    n = xs.numel()
    cam = (ctypes.c_double * 9)(*[float(v) for v in camera.reshape(9).tolist()])
    out = torch.empty(n, 3, dtype=torch.float32, device=dev)
    lib = _lib.load()
    _lib.check(lib.oryon_lift_pcd(_lib.handle(dev.index), ptr(d), _DEPTH_DTYPES[d.dtype], H, W, cam, ptr(xs), ptr(ys), n, ptr(out),
                                  stream_ptr(dev)))
    return out


def corrs_to_pcd(corrs: Tensor, depth_a: Tensor, depth_q: Tensor, camera_a: Tensor, camera_q: Tensor,
                 featmap_size: Sequence[int], size_a: Sequence[int], size_q: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """Fused form of reference pipeline.py:447-460: scale featmap-space correspondences to raw-frame pixels,
    drop out-of-bounds rows, truncate, lift both views and convert to metres -> ``(pcd_a, pcd_q) [m,3]``."""
    dev = device_of(depth_a, depth_q, corrs)
    c = as_device(corrs, dev, torch.int64)
    da, dq = depth_a.squeeze(), depth_q.squeeze()
    if da.dtype != dq.dtype or da.dtype not in _DEPTH_DTYPES:
        da, dq = da.to(torch.float32), dq.to(torch.float32)
    da, dq = as_device(da, dev), as_device(dq, dev)
    n = c.shape[0]
    ka = (ctypes.c_double * 9)(*[float(v) for v in camera_a.reshape(9).tolist()])
    kq = (ctypes.c_double * 9)(*[float(v) for v in camera_q.reshape(9).tolist()])
    pa = torch.empty(n, 3, dtype=torch.float32, device=dev)
    pq = torch.empty(n, 3, dtype=torch.float32, device=dev)
    nv = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.oryon_corrs_to_pcd(_lib.handle(dev.index), ptr(c), n, int(featmap_size[0]), int(featmap_size[1]), ptr(da), ptr(dq),
                                      _DEPTH_DTYPES[da.dtype], int(size_a[0]), int(size_a[1]), int(size_q[0]), int(size_q[1]), ka, kq,
                                      ptr(pa), ptr(pq), ptr(nv), stream_ptr(dev)))
    m = int(nv.item())
    return pa[:m], pq[:m]


def select_lift_batched(rows: Tensor, roi_a: Tensor, roi_q: Tensor, nn_idx: Tensor, depth_a: Tensor, depth_q: Tensor,
                        cameras_a: Tensor, cameras_q: Tensor, featmap_size: Sequence[int]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """``oryon_select_lift``: for B pairs in one launch, the row selection at the end of ``nn_correspondences``
    (reference utils/pcd.py:207-212) and the scaling / bounds test / lifting of pipeline.py:447-460.

    ``rows int32 [B,n]``: positions in each pair's anchor ROI list chosen by the caller's draws (``rows[b,0] < 0`` for a
    pair without correspondences); ``roi_a/roi_q/nn_idx``: what went into / came out of ``match_nn``; ``depth_x
    [B,H,W]`` raw frames (mm); ``cameras_x [B,3,3]``.  Returns ``(corrs int64 [B,n,4], pcd_a [B,n,3], pcd_q [B,n,3],
    n_valid int32 [B])`` on the GPU; rows ``>= n_valid[b]`` of the clouds are undefined."""
    dev = device_of(roi_a, nn_idx)
    rows = as_device(rows, dev, torch.int32)
    B, n = rows.shape
    if depth_a.dtype != depth_q.dtype or depth_a.dtype not in _DEPTH_DTYPES:
        depth_a, depth_q = depth_a.to(torch.float32), depth_q.to(torch.float32)
    da, dq = as_device(depth_a, dev), as_device(depth_q, dev)
    if da.dim() != 3 or dq.dim() != 3 or da.shape[0] != B or dq.shape[0] != B:
        raise ValueError("select_lift_batched: depth frames must be [B,H,W]")
    ka = cameras_a.detach().to("cpu", torch.float64).reshape(B, 9).contiguous()
    kq = cameras_q.detach().to("cpu", torch.float64).reshape(B, 9).contiguous()
    corrs = torch.empty(B, n, 4, dtype=torch.int64, device=dev)
    pa = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    pq = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    nv = torch.empty(B, dtype=torch.int32, device=dev)
    c_dbl = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.load().oryon_select_lift(
        _lib.handle(dev.index), ptr(rows), B, n, ptr(roi_a), ptr(roi_q), ptr(nn_idx), roi_a.shape[1], roi_q.shape[1],
        int(featmap_size[0]), int(featmap_size[1]), ptr(da), ptr(dq), _DEPTH_DTYPES[da.dtype], da.shape[1], da.shape[2], dq.shape[1],
        dq.shape[2], ctypes.cast(ka.data_ptr(), c_dbl), ctypes.cast(kq.data_ptr(), c_dbl), ptr(corrs), ptr(pa), ptr(pq), ptr(nv),
        stream_ptr(dev)))
    return corrs, pa, pq, nv
