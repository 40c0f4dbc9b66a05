"""Seeded synthetic inputs shared by tests, bench.py and oracle/make_golden.py.

No dataset or checkpoint is available offline, so every parity case and
every benchmark runs on inputs generated here from an integer seed with the CPU generator (the
build container and the GPU box run the same torch build, so the streams are identical; the
golden fixtures additionally store a checksum of the generated inputs).
"""
from __future__ import annotations

import math

import numpy as np
from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

NOCS_INTRINSICS = (591.0125, 0.0, 322.525, 0.0, 590.16775, 244.11084, 0.0, 0.0, 1.0)  # reference datasets.py:398


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def ellipse_mask(h: int, w: int, frac: float, cy: float = 0.5, cx: float = 0.5, dtype=torch.int32) -> Tensor:
    """Centred axis-aligned ellipse covering about ``frac`` of the ``h x w`` grid (values 0/1)."""
    ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h - cy
    xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w - cx
    r2 = frac / math.pi  # area of ellipse with semi-axes a=b=sqrt(frac/pi) in unit square
    m = (ys[:, None] ** 2 + xs[None, :] ** 2) <= r2
    return m.to(dtype)


def smooth_feature_pair(seed: int, d: int, h: int, w: int, *, coarse: int = 8, shift: Tuple[int, int] = (3, -2),
                        noise: float = 0.05) -> Tuple[Tensor, Tensor]:
    """Two ``[D,H,W]`` float32 maps that look like decoder output: a bilinear up-sampling of a coarse
    random grid (neighbouring pixels are similar, so near-ties occur) and, for the query, the same
    map rolled by ``shift`` plus white noise (so true matches exist)."""
    g = _gen(seed)
    ch, cw = max(2, h // coarse), max(2, w // coarse)
    base = torch.randn(1, d, ch, cw, generator=g)
    fa = F.interpolate(base, size=(h, w), mode="bilinear", align_corners=True)[0]
    fa = fa + 0.02 * torch.randn(d, h, w, generator=g)
    fq = torch.roll(fa, shifts=shift, dims=(1, 2)) + noise * torch.randn(d, h, w, generator=g)
    return fa.contiguous(), fq.contiguous()


def permuted_feature_batch(seed: int, b: int, d: int, h: int, w: int, noise: float = 0.1,
                           device: str = "cpu", dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """BASELINE config 2 / 5 inputs: ``feat_a ~ N(0,1)`` of shape ``[B,D,H,W]``
    and ``feat_q = perm(feat_a) + noise*N(0,1)`` with a per-pair random pixel permutation, so each
    anchor pixel has exactly one strong match.  Returns ``(feat_a, feat_q, perm)`` where
    ``feat_q[b,:,perm[b,i]] ~ feat_a[b,:,i]`` (flattened pixels).  Generated on ``device`` with a
    device generator when it is CUDA (the bench needs 629 MB per batch; values are not golden)."""
    if device == "cpu":
        g = _gen(seed)
        fa = torch.randn(b, d, h * w, generator=g)
        perm = torch.stack([torch.randperm(h * w, generator=g) for _ in range(b)])
        nz = torch.randn(b, d, h * w, generator=g)
    else:
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        fa = torch.randn(b, d, h * w, generator=g, device=device)
        perm = torch.stack([torch.randperm(h * w, generator=g, device=device) for _ in range(b)])
        nz = torch.randn(b, d, h * w, generator=g, device=device)
    fq = torch.empty_like(fa)
    fq.scatter_(2, perm[:, None, :].expand(b, d, h * w), fa)
    fq.add_(nz, alpha=noise)
    return fa.view(b, d, h, w).to(dtype), fq.view(b, d, h, w).to(dtype), perm


def random_rotation(g: torch.Generator, max_angle_deg: float = 45.0) -> Tensor:
    axis = torch.randn(3, generator=g, dtype=torch.float64)
    axis = axis / axis.norm()
    ang = (torch.rand(1, generator=g, dtype=torch.float64).item() * 2 - 1) * math.radians(max_angle_deg)
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def rigid_correspondences(seed: int, n: int = 500, outlier_frac: float = 0.3, noise: float = 0.002,
                          extent: float = 0.3) -> Dict[str, Tensor]:
    """``n`` 3-D correspondences (metres) related by a planted rigid motion, with Gaussian noise on the
    inliers and ``outlier_frac`` of the targets replaced by uniform points: PointDSC parity input."""
    g = _gen(seed)
    src = (torch.rand(n, 3, generator=g, dtype=torch.float64) - 0.5) * extent
    src[:, 2] += 1.0
    R = random_rotation(g)
    t = (torch.rand(3, generator=g, dtype=torch.float64) - 0.5) * 0.2
    tgt = src @ R.T + t + noise * torch.randn(n, 3, generator=g, dtype=torch.float64)
    n_out = int(n * outlier_frac)
    out_idx = torch.randperm(n, generator=g)[:n_out]
    tgt[out_idx] = (torch.rand(n_out, 3, generator=g, dtype=torch.float64) - 0.5) * extent * 1.5 + torch.tensor([0, 0, 1.0], dtype=torch.float64) + t
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3], T[:3, 3] = R, t
    return dict(src=src.float(), tgt=tgt.float(), T=T.float(), outliers=out_idx)


POINTDSC_DEFAULT_CFG = dict(in_dim=6, num_layers=12, num_channels=128, num_iterations=10, ratio=0.1,
                            sigma_d=0.1, k=40, inlier_threshold=0.1)


POINTDSC_CASES = {300: (500, 0.3), 301: (500, 0.6), 302: (137, 0.2), 303: (41, 0.1), 304: (500, 0.0)}  # seed -> (n, outlier frac)


def pointdsc_state_dict(seed: int, cfg: Dict = POINTDSC_DEFAULT_CFG) -> Dict[str, Tensor]:
    """Seeded random PointDSC ``state_dict`` with the reference's parameter names and shapes
    (reference models/pointdsc/PointDSC.py:9-113): Xavier-normal conv weights, small random biases,
    *non-trivial* BatchNorm running statistics so that eval-mode BN folding is exercised."""
    g = _gen(seed)
    C, L = cfg["num_channels"], cfg["num_layers"]
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin):
        std = math.sqrt(2.0 / (cin + cout))
        sd[name + ".weight"] = torch.randn(cout, cin, 1, generator=g) * std
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.05 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    sd["sigma"] = torch.tensor([1.0])
    sd["sigma_spat"] = torch.tensor([float(cfg["sigma_d"])])
    conv("encoder.layer0", C, cfg["in_dim"])
    for i in range(L):
        p = f"encoder.blocks.PointCN_layer_{i}"
        conv(p + ".0", C, C)
        bn(p + ".1", C)
        p = f"encoder.blocks.NonLocal_layer_{i}"
        conv(p + ".fc_message.0", C // 2, C)
        bn(p + ".fc_message.1", C // 2)
        conv(p + ".fc_message.3", C // 2, C // 2)
        bn(p + ".fc_message.4", C // 2)
        conv(p + ".fc_message.6", C, C // 2)
        conv(p + ".projection_q", C, C)
        conv(p + ".projection_k", C, C)
        conv(p + ".projection_v", C, C)
    conv("classification.0", 32, C)
    conv("classification.2", 32, 32)
    conv("classification.4", 1, 32)
    return sd


def synthetic_rgbd_pair(seed: int, raw_hw: Tuple[int, int] = (480, 640)) -> Dict[str, Tensor]:
    """One synthetic NOCS-like RGB-D pair: a smooth depth surface (mm, int32)
    for the anchor; the query depth is produced by moving the anchor cloud rigidly and re-rendering it
    by nearest-pixel splatting (holes filled with the median), so a true relative pose exists."""
    g = _gen(seed)
    H, W = raw_hw
    K = torch.tensor(NOCS_INTRINSICS, dtype=torch.float64).view(3, 3)
    coarse = torch.rand(1, 1, 6, 8, generator=g) * 400.0 + 800.0
    depth_a = F.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=True)[0, 0].round().to(torch.int32)
    rgb_a = torch.rand(3, 224, 224, generator=g)
    rgb_q = torch.rand(3, 224, 224, generator=g)
    R = random_rotation(g, 10.0)
    t = (torch.rand(3, generator=g, dtype=torch.float64) - 0.5) * 0.05
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    z = depth_a.double() / 1000.0
    pts = torch.stack(((xs - K[0, 2]) * z / K[0, 0], (ys - K[1, 2]) * z / K[1, 1], z), -1).view(-1, 3)
    moved = pts @ R.T + t
    u = (moved[:, 0] / moved[:, 2] * K[0, 0] + K[0, 2]).round().long()
    v = (moved[:, 1] / moved[:, 2] * K[1, 1] + K[1, 2]).round().long()
    ok = (u >= 0) & (u < W) & (v >= 0) & (v < H)
    depth_q = torch.zeros(H * W, dtype=torch.float64)
    depth_q[v[ok] * W + u[ok]] = moved[ok, 2] * 1000.0
    med = depth_q[depth_q > 0].median()
    depth_q[depth_q == 0] = med
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3], T[:3, 3] = R, t
    return dict(rgb_a=rgb_a, rgb_q=rgb_q, depth_a=depth_a, depth_q=depth_q.view(H, W).round().to(torch.int32),
                camera=K, T_rel=T)


# name -> (seed, D, H, W, mask frac a, mask frac q, noise, threshold, max_corrs, subsample_source)
MATCH_CASES = {
    "m0_d32_48x48": (100, 32, 48, 48, 0.30, 0.35, 0.05, 0.25, 500, None),
    "m1_d32_64x64_subsample": (101, 32, 64, 64, 0.40, 0.45, 0.05, 0.25, 500, 300),
    "m2_d128_40x40_full": (102, 128, 40, 40, 2.0, 2.0, 0.05, 0.25, 500, 5000),
    "m3_d32_none": (103, 32, 32, 32, 0.25, 0.25, 50.0, 0.02, 500, None),
    "m4_d32_replacement": (104, 32, 32, 32, 0.05, 0.30, 0.05, 0.25, 500, None),
    "m5_d17_24x40_ragged": (105, 17, 24, 40, 0.30, 0.50, 0.05, 0.25, 64, 100),
    "m6_d32_96x96_refsize": (106, 32, 96, 96, 0.15, 0.20, 0.08, 0.25, 500, 5000),
}


def match_inputs(case):
    seed, d, h, w, fa_frac, fq_frac, noise, th, max_corrs, sub = MATCH_CASES[case]
    fa, fq = smooth_feature_pair(seed, d, h, w, noise=noise)
    ma = ellipse_mask(h, w, fa_frac, 0.5, 0.5)
    mq = ellipse_mask(h, w, fq_frac, 0.55, 0.45)
    return fa, fq, ma, mq, th, max_corrs, sub, seed



LIFT_CASES = {200: ((192, 192), (480, 640)), 201: ((192, 192), (200, 150)), 202: ((48, 64), (480, 640))}


def lift_inputs(seed: int):
    """Random featmap-space correspondences + integer depth maps (10 % zero-depth pixels, which the
    reference does not filter: utils/pcd.py:52,72-74) for the scale/bounds/lift parity cases."""
    (HO, WO), (H, W) = LIFT_CASES[seed]
    g = _gen(seed)
    corrs = torch.stack([torch.randint(0, HO, (500,), generator=g), torch.randint(0, WO, (500,), generator=g),
                         torch.randint(0, HO, (500,), generator=g), torch.randint(0, WO, (500,), generator=g)], 1)
    depth_a = torch.randint(0, 3000, (H, W), generator=g, dtype=torch.int32)
    depth_q = torch.randint(0, 3000, (H, W), generator=g, dtype=torch.int32)
    depth_a[torch.rand(H, W, generator=g) < 0.1] = 0
    K = torch.tensor(NOCS_INTRINSICS, dtype=torch.float64)
    return corrs, depth_a, depth_q, K, (HO, WO), (H, W)


def tensor_checksum(t: Tensor) -> str:
    """sha256 of the raw bytes: pins regenerated inputs to the ones the golden outputs came from."""
    import hashlib
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def synthetic_batch(seed: int, B: int, *, empty_mask_pairs=(), mask_frac: float = 0.12) -> Dict:
    """A batch dict with the schema of the reference's ``CollateWrapper.__call__`` (datasets.py:202-245), filled
    with synthetic RGB-D pairs (``synthetic_rgbd_pair``): the input contract of ``FPM_Pipeline.test_step``.
    ``prompt_tokens [B,80,77]`` stands in for ``prompt`` (no BPE vocabulary offline)."""
    from . import synth_backbone
    pairs = [synthetic_rgbd_pair(seed * 1000 + i) for i in range(B)]
    masks_a = torch.stack([ellipse_mask(224, 224, mask_frac, 0.5, 0.5, torch.uint8) for _ in range(B)])
    masks_q = torch.stack([ellipse_mask(224, 224, mask_frac * 1.2, 0.52, 0.47, torch.uint8) for _ in range(B)])
    for i in empty_mask_pairs:
        masks_q[i] = 0
    K = torch.stack([p["camera"] for p in pairs])
    g = _gen(seed)
    pose_a = torch.eye(4, dtype=torch.float64).repeat(B, 1, 1)
    for i in range(B):
        pose_a[i, :3, :3] = random_rotation(g, 30.0)
        pose_a[i, :3, 3] = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64) + 0.1 * torch.randn(3, generator=g, dtype=torch.float64)
    pose_q = torch.stack([p["T_rel"] for p in pairs]) @ pose_a
    sizes = torch.tensor([[480, 640]] * B)

    def view(key_rgb, key_depth, masks, poses, tag):
        return dict(rgb=torch.stack([p[key_rgb] for p in pairs]), mask=masks, orig_depth=[p[key_depth] for p in pairs],
                    camera=K.clone(), pose=poses, sizes=sizes.clone(), instance_id=[f"{tag}_{seed}_{i}" for i in range(B)])

    return dict(anchor=view("rgb_a", "depth_a", masks_a, pose_a, "a"), query=view("rgb_q", "depth_q", masks_q, pose_q, "q"),
                prompt_tokens=synth_backbone.synthetic_tokens(seed, 1).expand(B, -1, -1).contiguous(),
                instance_id=[f"pair_{seed}_{i}" for i in range(B)], cls_id=[1] * B)


def planted_network_outputs(seed: int, B: int, shift=(2, 3), noise: float = 0.02):
    """Network outputs + batch for which the post-network path has a well-posed answer: the query feature map is
    the anchor's rolled by ``shift`` feature-map pixels (+ noise), the depth maps are near-planar and rolled by the
    same shift in raw pixels ((480/192, 640/192) per feature-map pixel), so true matches lift to 3-D points related
    by a (near) pure translation that PointDSC recovers.  Returns ``(outputs, batch)`` with CPU tensors."""
    g = _gen(seed)
    feats_a, feats_q, logit_a, logit_q = [], [], [], []
    batch = synthetic_batch(seed, B)
    for i in range(B):
        fa, _ = smooth_feature_pair(seed * 100 + i, 32, 192, 192, coarse=6, noise=noise)
        fq = torch.roll(fa, shifts=shift, dims=(1, 2)) + noise * torch.randn(32, 192, 192, generator=g)
        ma = ellipse_mask(192, 192, 0.10 + 0.03 * i, 0.5, 0.5, torch.float32)
        mq = torch.roll(ma, shifts=shift, dims=(0, 1))
        feats_a.append(fa), feats_q.append(fq)
        logit_a.append((ma * 2 - 1) * (0.5 + torch.rand(192, 192, generator=g)))
        logit_q.append((mq * 2 - 1) * (0.5 + torch.rand(192, 192, generator=g)))
        coarse = torch.rand(1, 1, 4, 5, generator=g) * 60.0 + 970.0
        da = F.interpolate(coarse, size=(480, 640), mode="bicubic", align_corners=True)[0, 0].round().to(torch.int32)
        dq = torch.roll(da, shifts=(int(shift[0] * 480 / 192), int(shift[1] * 640 / 192)), dims=(0, 1))
        batch["anchor"]["orig_depth"][i], batch["query"]["orig_depth"][i] = da, dq
    outputs = dict(featmap_a=torch.stack(feats_a), featmap_q=torch.stack(feats_q),
                   mask_a=torch.stack(logit_a).unsqueeze(1), mask_q=torch.stack(logit_q).unsqueeze(1))
    return outputs, batch


# ------------------------------------------------------------------------------------------------
# evaluator inputs: synthetic object models / symmetry sets / pose pairs
# ------------------------------------------------------------------------------------------------
def _axis_rotation(axis, angle: float) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def eval_objects(seed: int = 0) -> Dict:
    """Three synthetic objects in the format of the reference datasets' ``get_object_info()`` (datasets.py:509-513):
    ``models[id]['pts'] [n,3] float64`` (mm), ``diams[id]`` (mm), ``syms[id]`` = list of ``{'R': [3,3], 't': [3,1]}``.
    1: asymmetric, 2: continuous symmetry about z discretised in 36 steps, 3: one discrete 180-degree symmetry
    with an offset."""
    rng = np.random.RandomState(seed)
    models, diams, syms = {}, {}, {}
    pts = (rng.rand(700, 3) - 0.5) * np.array([60.0, 90.0, 120.0])
    models[1], syms[1] = dict(pts=pts), [dict(R=np.eye(3), t=np.zeros((3, 1)))]
    ang, h = rng.rand(1500) * 2 * np.pi, (rng.rand(1500) - 0.5) * 150.0
    models[2] = dict(pts=np.stack([40.0 * np.cos(ang), 40.0 * np.sin(ang), h], axis=1))
    syms[2] = [dict(R=_axis_rotation([0, 0, 1], 2 * np.pi * i / 36), t=np.zeros((3, 1))) for i in range(36)]
    pts = (rng.rand(300, 3) - 0.5) * np.array([200.0, 50.0, 80.0]) + np.array([5.0, 0.0, 0.0])
    R180 = _axis_rotation([0, 1, 0], np.pi)
    off = np.array([[5.0], [0.0], [0.0]])
    models[3] = dict(pts=pts)
    syms[3] = [dict(R=np.eye(3), t=np.zeros((3, 1))), dict(R=R180, t=-R180 @ off + off)]
    for k, m in models.items():
        p = m["pts"]
        diams[k] = float(np.sqrt(((p[:, None, :] - p[None, ::7, :]) ** 2).sum(-1)).max())
    return dict(models=models, diams=diams, syms=syms)


def eval_cases(seed: int = 0, n: int = 18) -> Dict:
    """``n`` (prediction, ground truth) pose pairs in the shapes ``test_step`` hands to ``Evaluator.register_test``
    (pipeline.py:321-331): ``pred_pose`` / ``pred_pose_rel`` float32 ``[n,4,4]``, ``gt_pose`` float64, metres; errors
    from a fraction of a degree / millimetre up to gross failures; entry 4 is an identity ``pred_pose_rel`` (failed
    registration), entry 5 an all-zero one."""
    g = _gen(7000 + seed)
    pose_a, gt, pred, rel, cls = [], [], [], [], []
    for i in range(n):
        Ta, Tq = torch.eye(4, dtype=torch.float64), torch.eye(4, dtype=torch.float64)
        Ta[:3, :3], Tq[:3, :3] = random_rotation(g, 180.0), random_rotation(g, 180.0)
        Ta[:3, 3] = torch.tensor([0.0, 0.0, 0.9], dtype=torch.float64) + 0.15 * torch.randn(3, generator=g, dtype=torch.float64)
        Tq[:3, 3] = torch.tensor([0.0, 0.0, 0.8], dtype=torch.float64) + 0.15 * torch.randn(3, generator=g, dtype=torch.float64)
        scale = [0.2, 1.0, 3.0, 8.0, 20.0, 60.0][i % 6]
        E = torch.eye(4, dtype=torch.float64)
        E[:3, :3] = random_rotation(g, scale)
        E[:3, 3] = 0.002 * scale * torch.randn(3, generator=g, dtype=torch.float64)
        Trel = (E @ Tq @ torch.linalg.inv(Ta)).to(torch.float32)
        if i == 4:
            Trel = torch.eye(4)
        if i == 5:
            Trel = torch.zeros(4, 4)
        pose_a.append(Ta), gt.append(Tq), rel.append(Trel), pred.append(Trel @ Ta.to(torch.float32)), cls.append(1 + i % 3)
    K = torch.tensor(NOCS_INTRINSICS, dtype=torch.float64).reshape(3, 3)
    return dict(pose_a=torch.stack(pose_a), gt_pose=torch.stack(gt), pred_pose=torch.stack(pred), pred_pose_rel=torch.stack(rel),
                cls_id=cls, camera=K, iou_a=torch.rand(n, generator=g), iou_q=torch.rand(n, generator=g),
                instance_id=[f"scene_{i}" for i in range(n)])


def raw_frames(seed: int, n: int, hw: Tuple[int, int] = (480, 640)):
    """``n`` decoded frames as the dataset readers return them (utils/data/nocs.py ``get_item_data``): ``rgb`` uint8 HWC,
    ``mask`` uint8 label image (several instances, one of them ``mask_id``), ``depth`` int32 mm."""
    rng = np.random.RandomState(9000 + seed)
    H, W = hw
    out = []
    for i in range(n):
        yy, xx = np.mgrid[0:H, 0:W]
        base = np.stack([(xx * 255 // W), (yy * 255 // H), ((xx + yy) * 255 // (H + W))], axis=2).astype(np.int32)
        rgb = np.clip(base + rng.randint(-60, 60, size=(H, W, 3)), 0, 255).astype(np.uint8)
        mask = np.zeros((H, W), np.uint8)
        for lab, (cy, cx, ry, rx) in enumerate([(0.3, 0.3, 0.12, 0.1), (0.55, 0.6, 0.2, 0.17), (0.8, 0.2, 0.08, 0.15)], start=3):
            mask[((yy - cy * H) / (ry * H)) ** 2 + ((xx - cx * W) / (rx * W)) ** 2 <= 1.0] = lab
        depth = rng.randint(400, 1800, size=(H, W)).astype(np.int32)
        out.append(dict(rgb=rgb, mask=mask, depth=depth, mask_id=4, instance_id=f"frame_{seed}_{i}"))
    return out


def eval_mesh_objects(seed: int = 0) -> Dict:
    """Three closed triangle meshes (``pts`` mm, ``faces``) in the ``get_object_info()`` format, for the VSD / AR part of the
    evaluator: 1 an asymmetric ellipsoid with a bump, 2 a 24-gon prism (24-fold symmetry about z), 3 a box (2-fold)."""
    rng = np.random.RandomState(100 + seed)
    models, diams, syms = {}, {}, {}

    def lathe(profile_r, profile_z, n_seg, scale=(1.0, 1.0, 1.0), bump=None):
        rings = len(profile_r)
        ang = np.arange(n_seg) * 2 * np.pi / n_seg
        pts = [[0.0, 0.0, profile_z[0] - 1e-9]]
        for r_, z_ in zip(profile_r, profile_z):
            for a_ in ang:
                pts.append([r_ * np.cos(a_), r_ * np.sin(a_), z_])
        pts.append([0.0, 0.0, profile_z[-1] + 1e-9])
        pts = np.asarray(pts) * np.asarray(scale)
        if bump is not None:
            d = np.linalg.norm(pts - bump[:3], axis=1)
            pts = pts * (1.0 + bump[3] * np.exp(-(d / bump[4]) ** 2))[:, None]
        faces = []
        ring = lambda k, i: 1 + k * n_seg + (i % n_seg)  # noqa: E731
        for i in range(n_seg):
            faces.append([0, ring(0, i + 1), ring(0, i)])
            faces.append([len(pts) - 1, ring(rings - 1, i), ring(rings - 1, i + 1)])
            for k in range(rings - 1):
                faces.append([ring(k, i), ring(k, i + 1), ring(k + 1, i + 1)])
                faces.append([ring(k, i), ring(k + 1, i + 1), ring(k + 1, i)])
        return pts, np.asarray(faces, dtype=np.int64)

    th = np.linspace(0.15, np.pi - 0.15, 14)
    p, f = lathe(50 * np.sin(th), -50 * np.cos(th), 28, scale=(1.0, 0.7, 1.3), bump=np.array([40.0, 10.0, 20.0, 0.35, 30.0]))
    models[1], syms[1] = dict(pts=p, faces=f), [dict(R=np.eye(3), t=np.zeros((3, 1)))]
    p, f = lathe(np.array([45.0, 45.0, 30.0, 30.0]), np.array([-60.0, 10.0, 20.0, 60.0]), 24)
    models[2] = dict(pts=p, faces=f)
    syms[2] = [dict(R=_axis_rotation([0, 0, 1], 2 * np.pi * i / 24), t=np.zeros((3, 1))) for i in range(24)]
    p, f = lathe(np.array([60.0, 60.0]) * np.sqrt(2), np.array([-25.0, 25.0]), 4, scale=(1.0, 0.5, 1.0))
    p = p @ _axis_rotation([0, 0, 1], np.pi / 4).T
    models[3] = dict(pts=p, faces=f)
    syms[3] = [dict(R=np.eye(3), t=np.zeros((3, 1))), dict(R=_axis_rotation([0, 0, 1], np.pi), t=np.zeros((3, 1)))]
    for k, m in models.items():
        q = m["pts"]
        diams[k] = float(np.sqrt(((q[:, None, :] - q[None, :, :]) ** 2).sum(-1)).max())
    del rng
    return dict(models=models, diams=diams, syms=syms)


def eval_scene_depth(models: Dict, cls_id: int, gt_pose: np.ndarray, K: np.ndarray, seed: int, hw: Tuple[int, int] = (480, 640),
                     render=None) -> np.ndarray:
    """A test depth image (int32 mm) for the VSD evaluation of one pair: the object rendered in its ground-truth pose over a
    far background plane, an occluder strip across part of it and a band of missing (zero) depth.  ``render`` is the
    rasteriser to use (tests pass the oracle's; nothing in the product path calls this)."""
    H, W = hw
    rng = np.random.RandomState(500 + seed)
    m = models[cls_id]
    d = render(m["pts"], m["faces"], gt_pose[:3, :3], gt_pose[:3, 3] * 1000.0, K[0, 0], K[1, 1], K[0, 2], K[1, 2], H, W)
    scene = np.where(d > 0, d + rng.randn(H, W).astype(np.float32) * 1.0, 1800.0 + rng.randn(H, W).astype(np.float32) * 2.0)
    ys, xs = np.nonzero(d > 0)
    if ys.size:
        yc, xc = int(ys.mean()), int(xs.mean())
        scene[max(0, yc - 4):yc + 5, :] = np.minimum(scene[max(0, yc - 4):yc + 5, :], float(d[d > 0].min()) - 60.0)   # occluder
        scene[:, max(0, xc + 6):xc + 12] = 0.0                                                                          # missing depth
    return np.round(scene).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# a tiny dataset tree in the on-disk layout of the reference's NOCS (REAL275) test split
# (datasets.py:369-457, utils/data/nocs.py): input of the reader tests and of oracle/make_golden_nocs.py
# ------------------------------------------------------------------------------------------------
NOCS_TREE_OBJECTS = {"mug_synth_a": (6, ["mug", "white", "red"]), "bowl_synth_b": (2, ["bowl", "blue", "green"]),
                     "can_synth_c": (3, ["can", "metal", "paper"])}   # object name -> (category id, [class name, descriptions...])


def write_nocs_tree(root: str, seed: int = 0, name: str = "nocs", split: str = "cross_scene_test", hw: Tuple[int, int] = (48, 64),
                    n_scenes: int = 2, n_imgs: int = 3, mask_scale: int = 1) -> Dict:
    """Writes ``<root>/<name>/...`` with ``n_scenes x n_imgs`` frames (colour / label mask / 16-bit depth PNGs, meta and
    detection text files, per-frame pose pickles), three object models (one with a continuous, one with a discrete
    symmetry), the prompt templates, the object split and a fixed pair split (instance list, annotations, tracked list).
    Deterministic in ``seed``.  One frame lacks the third object, one object is a single pixel, one pair has no ground-truth
    correspondences (an invalid sample in the reference's sense)."""
    import json
    import os
    import pickle
    from PIL import Image

    g = np.random.default_rng(seed)
    base = os.path.join(root, name)
    H, W = hw
    names = list(NOCS_TREE_OBJECTS)
    for d in ("gts/real_test", "obj_models/real_test", f"fixed_split/{split}", "split/real_test"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    with open(os.path.join(base, "templates.json"), "w") as f:
        json.dump(["a photo of a {}.", "a bad photo of the {}.", "a close-up photo of a {}.", "itap of a {}.", "a {} in a scene."], f)
    with open(os.path.join(base, "object_splits.json"), "w") as f:
        json.dump({"all": ["6", "2", "3"], "mugs": ["6"]}, f)
    with open(os.path.join(base, "obj_names.json"), "w") as f:
        json.dump({k: v[1] + ["norm"] for k, v in NOCS_TREE_OBJECTS.items()}, f)

    # object models: vertices in metres (the reader scales to mm), normals, OBJ faces (1-based, v/vt/vn triplets)
    info = {}
    for i, obj in enumerate(names):
        nv = 12 + 4 * i
        v = g.normal(size=(nv, 3)) * 0.05
        nrm = v / np.linalg.norm(v, axis=1, keepdims=True)
        faces = np.stack([np.arange(nv - 2), np.arange(1, nv - 1), np.arange(2, nv)], 1) + 1
        p = os.path.join(base, "obj_models/real_test", obj)
        np.savetxt(p + "_vertices.txt", v, fmt="%.6f")
        np.savetxt(p + "_normals.txt", nrm, fmt="%.6f")
        with open(p + ".obj", "w") as f:
            f.write("# synthetic\n" + "".join(f"v {a:.6f} {b:.6f} {c:.6f}\n" for a, b, c in v))
            f.write("".join(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n" for a, b, c in faces))
        info[obj] = {"diameter": float(np.ptp(v, axis=0).max() * 1000.0)}
    info[names[1]]["symmetries_continuous"] = [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]
    info[names[2]]["symmetries_discrete"] = [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]
    with open(os.path.join(base, "obj_models/real_test/models_info.json"), "w") as f:
        json.dump(info, f)

    frames = []
    for s in range(1, n_scenes + 1):
        os.makedirs(os.path.join(base, f"split/real_test/scene_{s}"), exist_ok=True)
        for im in range(n_imgs):
            present = names if not (s == 1 and im == 1) else names[:2]          # one frame without the third object
            rgb = g.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
            depth = g.integers(400, 1600, size=(H, W)).astype(np.uint16)
            mask = np.full((H, W), 255, np.uint8)
            meta, det, poses = [], [], []
            for k, obj in enumerate(present):
                mask_id = k + 1
                y0, x0 = (4 + 12 * k) * mask_scale, (5 + 14 * k) * mask_scale
                hh, ww = (10 * mask_scale, 12 * mask_scale) if not (s == 2 and im == 2 and k == 0) else (1, 1)   # one single-pixel object
                mask[y0:y0 + hh, x0:x0 + ww] = mask_id
                meta.append(f"{mask_id} {NOCS_TREE_OBJECTS[obj][0]} {obj}\n")
                det.append(f"{mask_id} {x0} {y0} {ww - 1} {hh - 1}\n")
                T = np.eye(4)
                T[:3, :3] = _axis_rotation(g.normal(size=3), float(g.uniform(0.1, 2.0))) * float(g.uniform(0.2, 0.4))   # NOCS poses carry a scale
                T[:3, 3] = g.normal(size=3) * 0.1 + np.array([0.0, 0.0, 1.0])
                poses.append(T)
            stem = os.path.join(base, f"split/real_test/scene_{s}/{im:04d}")
            Image.fromarray(rgb).save(stem + "_color.png")
            Image.fromarray(mask).save(stem + "_mask.png")
            Image.fromarray(depth).save(stem + "_depth.png")
            open(stem + "_meta.txt", "w").writelines(meta)
            open(stem + "_detection.txt", "w").writelines(det)
            with open(os.path.join(base, f"gts/real_test/results_real_test_scene_{s}_{im:04d}.pkl"), "wb") as f:
                pickle.dump({"gt_RTs": np.stack(poses)}, f)
            frames.append((s, im, present))

    # fixed pair split: anchors in scene 1, queries in scene 2 (plus one same-scene pair), objects present in both frames
    lines, annots = [], {}
    pairs = [(1, 0, 2, 0, names[0]), (1, 0, 2, 1, names[1]), (1, 2, 2, 2, names[2]), (1, 1, 2, 0, names[1]), (1, 2, 2, 2, names[0]),
             (2, 1, 2, 0, names[0])]
    for sa, ia, sq, iq, obj in pairs:
        cat = NOCS_TREE_OBJECTS[obj][0]
        lines.append(f"real_test, {sa} {ia}, {sq} {iq}, {cat} {obj}\n")
        gt = np.eye(4)
        gt[:3, :3] = _axis_rotation(g.normal(size=3), float(g.uniform(0.1, 1.0)))
        gt[:3, 3] = g.normal(size=3) * 100.0                                   # millimetres on disk
        n_c = int(g.integers(0, 30)) if (sa, ia, sq, iq, obj) != pairs[3] else 0   # one pair without ground-truth correspondences
        corrs = np.stack([g.integers(0, H, n_c), g.integers(0, W, n_c), g.integers(0, H, n_c), g.integers(0, W, n_c)], 1).astype(np.int64)
        annots[f"{sa}_{ia}_{sq}_{iq}_{cat}_{obj}"] = {"gt": gt, "corrs": corrs}
    sp = os.path.join(base, f"fixed_split/{split}")
    open(os.path.join(sp, "instance_list.txt"), "w").writelines(lines)
    open(os.path.join(sp, "tracked.txt"), "w").writelines(lines[:2])
    with open(os.path.join(sp, "annots.pkl"), "wb") as f:
        pickle.dump(annots, f)
    return dict(base=base, name=name, split=split, pairs=pairs, frames=frames)


def write_toyl_tree(root: str, seed: int = 0, name: str = "toyl", split: str = "cross_scene_test", hw: Tuple[int, int] = (48, 64),
                    n_scenes: int = 2, n_imgs: int = 3, mask_scale: int = 1) -> Dict:
    """Writes ``<root>/<name>/...`` in the on-disk layout of the reference's TOYL data (datasets.py:546-630,
    utils/data/toyl.py): BOP scene folders ``split/test/<scene:06d>/{rgb,depth,mask_visib}/<img:06d>.png`` with
    ``scene_gt.json`` / ``scene_gt_info.json``, ``models_name.json``, ``models_bop/obj_<id:06d>.ply`` (one ASCII, the others
    binary little endian) + ``models_info.json``, templates, object split and the fixed pair split.  Deterministic in ``seed``."""
    import json
    import os
    import pickle
    from PIL import Image

    g = np.random.default_rng(1000 + seed)
    base = os.path.join(root, name)
    H, W = hw
    objs = {1: ["duck", "yellow", "green"], 5: ["car", "red", "blue"], 12: ["robot", "grey", "pink"]}
    for d in ("models_bop", f"fixed_split/{split}"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    with open(os.path.join(base, "templates.json"), "w") as f:
        json.dump(["a photo of a {}.", "a blurry photo of the {}.", "a toy {}."], f)
    with open(os.path.join(base, "object_splits.json"), "w") as f:
        json.dump({"all": ["1", "5", "12"], "cars": ["5"], "ducks": ["1"]}, f)
    with open(os.path.join(base, "models_name.json"), "w") as f:
        json.dump({str(k): v for k, v in objs.items()}, f)

    info = {}
    for n, oid in enumerate(objs):
        nv = 10 + 3 * n
        v = (g.normal(size=(nv, 3)) * 40.0).astype(np.float32)
        nrm = (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
        faces = np.stack([np.arange(nv - 2), np.arange(1, nv - 1), np.arange(2, nv)], 1).astype(np.int32)
        header = (f"ply\nformat {'ascii' if n == 0 else 'binary_little_endian'} 1.0\ncomment synthetic\nelement vertex {nv}\n"
                  "property float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
                  f"element face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n")
        with open(os.path.join(base, "models_bop", f"obj_{oid:06d}.ply"), "wb") as f:
            f.write(header.encode())
            if n == 0:
                for a, b in zip(v, nrm):
                    f.write((" ".join(repr(float(x)) for x in (*a, *b)) + "\n").encode())
                for fc in faces:
                    f.write(("3 " + " ".join(str(int(x)) for x in fc) + "\n").encode())
            else:
                f.write(np.concatenate([v, nrm], 1).astype("<f4").tobytes())
                for fc in faces:
                    f.write(b"\x03" + fc.astype("<i4").tobytes())
        info[str(oid)] = {"diameter": float(np.ptp(v, axis=0).max())}
    info["5"]["symmetries_discrete"] = [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]
    info["12"]["symmetries_continuous"] = [{"axis": [0, 1, 0], "offset": [0, 0, 0]}]
    with open(os.path.join(base, "models_bop", "models_info.json"), "w") as f:
        json.dump(info, f)

    ids = list(objs)
    for s in range(1, n_scenes + 1):
        sd = os.path.join(base, "split", "test", f"{s:06d}")
        for d in ("rgb", "depth", "mask_visib"):
            os.makedirs(os.path.join(sd, d), exist_ok=True)
        gt, gt_info = {}, {}
        for im in range(n_imgs):
            present = ids if not (s == 2 and im == 1) else ids[1:]               # one frame without the first object
            Image.fromarray(g.integers(0, 256, size=(H, W, 3), dtype=np.uint8)).save(os.path.join(sd, "rgb", f"{im:06d}.png"))
            Image.fromarray(g.integers(300, 1400, size=(H, W)).astype(np.uint16)).save(os.path.join(sd, "depth", f"{im:06d}.png"))
            mask = np.zeros((H, W), np.uint8)
            gt[str(im)], gt_info[str(im)] = [], []
            for k, oid in enumerate(present):
                y0, x0, hh, ww = (3 + 13 * k) * mask_scale, (6 + 15 * k) * mask_scale, 9 * mask_scale, 11 * mask_scale
                mask[y0:y0 + hh, x0:x0 + ww] = k + 1
                R = _axis_rotation(g.normal(size=3), float(g.uniform(0.1, 2.5)))
                gt[str(im)].append({"cam_R_m2c": R.reshape(-1).tolist(), "cam_t_m2c": (g.normal(size=3) * 80.0 + np.array([0, 0, 900.0])).tolist(),
                                    "obj_id": oid})
                gt_info[str(im)].append({"bbox_visib": [x0, y0, ww, hh], "bbox_obj": [x0, y0, ww, hh], "px_count_visib": hh * ww})
            Image.fromarray(mask).save(os.path.join(sd, "mask_visib", f"{im:06d}.png"))
        with open(os.path.join(sd, "scene_gt.json"), "w") as f:
            json.dump(gt, f)
        with open(os.path.join(sd, "scene_gt_info.json"), "w") as f:
            json.dump(gt_info, f)

    pairs = [(1, 0, 2, 0, 1), (1, 1, 2, 2, 5), (1, 2, 2, 1, 12), (1, 0, 2, 2, 5), (2, 0, 2, 2, 1)]
    lines, annots = [], {}
    for n, (sa, ia, sq, iq, oid) in enumerate(pairs):
        lines.append(f"test, {sa} {ia}, {sq} {iq}, {oid}\n")
        pose = np.eye(4)
        pose[:3, :3] = _axis_rotation(g.normal(size=3), float(g.uniform(0.1, 1.0)))
        pose[:3, 3] = g.normal(size=3) * 120.0
        n_c = int(g.integers(1, 25)) if n != 3 else 0
        corrs = np.stack([g.integers(0, H, n_c), g.integers(0, W, n_c), g.integers(0, H, n_c), g.integers(0, W, n_c)], 1).astype(np.int64)
        annots[f"{sa}_{ia}_{sq}_{iq}_{oid}"] = {"gt": pose, "corrs": corrs}
    sp = os.path.join(base, f"fixed_split/{split}")
    open(os.path.join(sp, "instance_list.txt"), "w").writelines(lines)
    open(os.path.join(sp, "tracked.txt"), "w").writelines(lines[1:3])
    with open(os.path.join(sp, "annots.pkl"), "wb") as f:
        pickle.dump(annots, f)
    return dict(base=base, name=name, split=split, pairs=pairs)


# ------------------------------------------------------------------------------------------------
# textured frames for the key-point baselines (scripts/evaluation/sift_baseline.py)
# ------------------------------------------------------------------------------------------------

# case -> (frame size, seed, query related to anchor, distance threshold)
SIFT_CASES = {"k0_small": ((160, 200), 900, True, 0.25), "k1_large": ((240, 320), 901, True, 0.25), "k2_unrelated": ((120, 160), 902, False, 0.02),
              "k3_tight": ((160, 200), 903, True, 0.05)}


def _texture(g: np.random.Generator, h: int, w: int) -> np.ndarray:
    """Float image in [0,1] with structure at several scales (blobs + edges): box-filtered noise octaves and rectangles."""
    img = np.zeros((h, w))
    for k, amp in ((3, 0.35), (7, 0.3), (15, 0.2), (31, 0.15)):
        n = g.normal(size=(h + k, w + k))
        c = np.cumsum(np.cumsum(n, 0), 1)
        box = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / k          # variance-preserving box filter
        img += amp * box[:h, :w]
    img = (img - img.min()) / (img.max() - img.min())
    for _ in range(max(6, h * w // 6000)):
        y0, x0 = int(g.integers(0, h - 8)), int(g.integers(0, w - 8))
        hh, ww = int(g.integers(4, max(5, h // 6))), int(g.integers(4, max(5, w // 6)))
        img[y0:y0 + hh, x0:x0 + ww] = 0.6 * img[y0:y0 + hh, x0:x0 + ww] + 0.4 * g.uniform()
    return img


def _warped_view(tex: np.ndarray, g: np.random.Generator, H: int, W: int, pad: int) -> np.ndarray:
    """The texture seen through a small similarity transform (bilinear sampling) with additive noise."""
    ang, s = np.deg2rad(g.uniform(-6, 6)), g.uniform(0.95, 1.05)
    ty, tx = g.uniform(-8, 8, size=2)
    yy, xx = np.meshgrid(np.arange(H) - H / 2, np.arange(W) - W / 2, indexing="ij")
    sy = s * (np.cos(ang) * yy - np.sin(ang) * xx) + H / 2 + pad + ty
    sx = s * (np.sin(ang) * yy + np.cos(ang) * xx) + W / 2 + pad + tx
    sy, sx = np.clip(sy, 0, tex.shape[0] - 1.001), np.clip(sx, 0, tex.shape[1] - 1.001)
    y0, x0 = np.floor(sy).astype(int), np.floor(sx).astype(int)
    fy, fx = sy - y0, sx - x0
    return (tex[y0, x0] * (1 - fy) * (1 - fx) + tex[y0, x0 + 1] * (1 - fy) * fx + tex[y0 + 1, x0] * fy * (1 - fx)
            + tex[y0 + 1, x0 + 1] * fy * fx) + g.normal(size=(H, W)) * 0.01


def _to_u8(a: np.ndarray) -> np.ndarray:
    return np.clip(np.round(a * 255.0), 0, 255).astype(np.uint8)


def textured_frame_pair(seed: int, hw: Tuple[int, int], related: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Two uint8 grey frames ``[H,W]``: the anchor is a crop of a synthetic texture, the query the same texture seen through a
    small similarity transform with additive noise -- or, with ``related=False``, another texture."""
    g = np.random.default_rng(seed)
    H, W = hw
    pad = 24
    tex = _texture(g, H + 2 * pad, W + 2 * pad)
    anchor = tex[pad:pad + H, pad:pad + W]
    query = _warped_view(tex, g, H, W, pad) if related else _texture(g, H, W)
    return _to_u8(anchor), _to_u8(query)


def textured_frames(seed: int, hw: Tuple[int, int], n: int) -> list:
    """``n`` uint8 grey views ``[H,W]`` of one synthetic texture (the first is the plain crop): any two of them match."""
    g = np.random.default_rng(seed)
    H, W = hw
    pad = 24
    tex = _texture(g, H + 2 * pad, W + 2 * pad)
    return [_to_u8(tex[pad:pad + H, pad:pad + W])] + [_to_u8(_warped_view(tex, g, H, W, pad)) for _ in range(n - 1)]


def rescale_cases(seed: int = 0) -> Dict:
    """name -> (coords, orig_scale, new_scale) for ``utils.misc.rescale_coords``: feature-map -> image and back, with out-of-range rows."""
    g = _gen(4000 + seed)
    f2 = torch.rand(40, 2, generator=g) * 230.0 - 10.0
    i4 = torch.randint(-5, 200, (37, 4), generator=g)
    fb = torch.rand(3, 25, 4, generator=g) * 192.0
    return {"float_2col_up": (f2, (192, 192), (480, 640)), "int_4col_up": (i4, (192, 192), (224, 224)),
            "float_batched_down": (fb, (192, 192), (48, 64)), "int_2col_down": (i4[:, :2].clone(), (192, 192), (24, 24))}
