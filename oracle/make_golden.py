"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference modules from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference has no tests or golden vectors of its own; these fixtures are the
pin for ``oracle/oryon_oracle.py`` and, through it, for the CUDA path.  Inputs are regenerated in the
tests from seeds by ``oryon_b200/synth.py``; each fixture stores a checksum of its inputs so that a
drift in the generator is detected rather than silently compared against stale outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shims  # noqa: E402

ref_shims.install()

from utils import coordinates as ref_coords  # noqa: E402  (reference)
from utils.pcd import lift_pcd as ref_lift_pcd  # noqa: E402
from utils.pcd import nn_correspondences as ref_nn_correspondences  # noqa: E402
from utils.pcd import pdist as ref_pdist  # noqa: E402
from models.pointdsc.PointDSC import PointDSC as RefPointDSC  # noqa: E402

from oryon_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MATCH_CASES = synth.MATCH_CASES
match_inputs = synth.match_inputs


def gen_match():
    for case in MATCH_CASES:
        fa, fq, ma, mq, th, max_corrs, sub, seed = match_inputs(case)
        # deterministic, pre-sampling quantities straight from the reference's distance function
        roi1, roi2 = torch.nonzero(ma == 1), torch.nonzero(mq == 1)
        f1 = fa[:, roi1[:, 0], roi1[:, 1]].T.float()
        f2 = fq[:, roi2[:, 0], roi2[:, 1]].T.float()
        dist = ref_pdist(f1, f2, "inv_norm_cosine")
        min_dist, nn_idx = torch.amin(dist, dim=1), torch.argmin(dist, dim=1)
        # top-2 margin in float64 (used by the CUDA parity rule: exact index where margin > tol)
        d64 = 0.5 * (1 - torch.nn.functional.normalize(f1.double(), dim=1) @ torch.nn.functional.normalize(f2.double(), dim=1).T)
        top2 = torch.topk(d64, k=2, dim=1, largest=False)[0]
        margin = (top2[:, 1] - top2[:, 0]).float()
        # the full reference function, RNG seeded the way the reference seeds it (utils/misc.py:186-196)
        torch.manual_seed(seed)
        corrs = ref_nn_correspondences(fa, fq, ma, mq, th, max_corrs, sub, "cpu")
        np.savez_compressed(
            os.path.join(OUT, f"match_{case}.npz"),
            in_sum=np.array([synth.tensor_checksum(fa), synth.tensor_checksum(fq),
                             synth.tensor_checksum(ma), synth.tensor_checksum(mq)]),
            n1=roi1.shape[0], n2=roi2.shape[0],
            min_dist=min_dist.numpy(), nn_idx=nn_idx.numpy().astype(np.int32), margin=margin.numpy(),
            corrs=(corrs.numpy() if corrs is not None else np.zeros((0, 4), np.int64)),
            is_none=np.array(corrs is None),
        )
        print(case, "n1", roi1.shape[0], "n2", roi2.shape[0], "valid", int((min_dist < th).sum()),
              "corrs", None if corrs is None else tuple(corrs.shape))


def gen_lift():
    """scale_coords + get_valid_coords + long-truncation + lift_pcd exactly as pipeline.py:447-460."""
    for seed in synth.LIFT_CASES:
        corrs, depth_a, depth_q, K, (HO, WO), (H, W) = synth.lift_inputs(seed)
        ca = ref_coords.scale_coords(corrs[:, :2].clone(), (HO, WO), (H, W))
        cq = ref_coords.scale_coords(corrs[:, 2:].clone(), (HO, WO), (H, W))
        valid = torch.logical_and(ref_coords.get_valid_coords(ca, (H, W)), ref_coords.get_valid_coords(cq, (H, W)))
        ca, cq = ca[valid].to(torch.long), cq[valid].to(torch.long)
        pa = ref_lift_pcd(depth_a.unsqueeze(-1), K, (ca[:, 1], ca[:, 0])) / 1000.0
        pq = ref_lift_pcd(depth_q.unsqueeze(-1), K, (cq[:, 1], cq[:, 0])) / 1000.0
        np.savez_compressed(os.path.join(OUT, f"lift_{seed}.npz"),
                            in_sum=np.array([synth.tensor_checksum(corrs), synth.tensor_checksum(depth_a), synth.tensor_checksum(depth_q)]),
                            ca=ca.numpy().astype(np.int32), cq=cq.numpy().astype(np.int32), pcd_a=pa.numpy(), pcd_q=pq.numpy())
        print("lift", seed, tuple(pa.shape), pa.dtype)


def gen_pointdsc():
    cfg = synth.POINTDSC_DEFAULT_CFG
    for seed, (n, out_frac) in synth.POINTDSC_CASES.items():
        sd = synth.pointdsc_state_dict(seed)
        model = RefPointDSC(in_dim=cfg["in_dim"], num_layers=cfg["num_layers"], num_channels=cfg["num_channels"],
                            num_iterations=cfg["num_iterations"], ratio=cfg["ratio"], sigma_d=cfg["sigma_d"],
                            k=cfg["k"], nms_radius=cfg["inlier_threshold"])
        missing = model.load_state_dict(sd, strict=True)
        model.eval()
        data = synth.rigid_correspondences(seed, n=n, outlier_frac=out_frac)
        pcd1, pcd2 = data["src"], data["tgt"]
        # body of reference utils/pointdsc/init.py:18-29 (module itself needs easydict at import only)
        corr_pos = torch.cat([pcd1, pcd2], axis=-1)
        corr_pos = corr_pos - corr_pos.mean(0)
        batch = {"corr_pos": corr_pos.unsqueeze(0).float(), "src_keypts": pcd1.unsqueeze(0).float(),
                 "tgt_keypts": pcd2.unsqueeze(0).float(), "testing": True}
        with torch.no_grad():
            res = model(batch)
            final = res["final_trans"].squeeze(0).float()
            # intermediates via the reference's own sub-modules
            src_dist = torch.norm(pcd1[None, :, None, :] - pcd1[None, None, :, :], dim=-1)
            comp = src_dist - torch.norm(pcd2[None, :, None, :] - pcd2[None, None, :, :], dim=-1)
            comp = torch.clamp(1.0 - comp ** 2 / model.sigma_spat ** 2, min=0)
            feat = model.encoder(batch["corr_pos"].permute(0, 2, 1), comp)
            conf = model.classification(feat).squeeze(1)
            seeds = model.pick_seeds(src_dist, conf, R=model.nms_radius, max_num=int(n * model.ratio))
        np.savez_compressed(os.path.join(OUT, f"pointdsc_{seed}.npz"), n=n, out_frac=out_frac,
                            in_sum=np.array([synth.tensor_checksum(pcd1), synth.tensor_checksum(pcd2),
                                             synth.tensor_checksum(sd["encoder.layer0.weight"])]),
                            final_trans=final.numpy(), planted=data["T"].numpy(), conf=conf[0].numpy(),
                            feat_sample=feat[0, :, :8].numpy(), seeds=seeds[0].numpy(), sc_sample=comp[0, :8, :8].numpy())
        err = (final - data["T"]).abs().max().item()
        print("pointdsc", seed, "n", n, "outliers", out_frac, "|T - planted|max", f"{err:.2e}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["match", "lift", "pointdsc"]
    torch.set_num_threads(8)
    if "match" in which:
        gen_match()
    if "lift" in which:
        gen_lift()
    if "pointdsc" in which:
        gen_pointdsc()
