"""Generate ``tests/golden/sift_kp_<case>.npz``: outputs of the reference's key-point matching functions
(``nn_correspondences`` of scripts/evaluation/sift_nocs.py:25-45 and sift_toyl.py:25-51, exec'd from the source text as it lies
under /root/reference, with the reference's ``utils.pcd.pdist`` and ``utils.misc.torch_sample_select`` underneath) on real
OpenCV SIFT descriptors of synthetic textured frames.

TEST INFRASTRUCTURE, build container only:  PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_sift.py

The descriptor sets are part of the fixture (SIFT descriptors are integer valued in [0,255]: stored as uint8; key points as
int16 (x, y) as the scripts cast them), so the tests do not depend on OpenCV.  Stored per case: the reference row argmin /
minimum / float64 top-2 margin of ``pdist(feats1, feats2, 'inv_norm_cosine')``, and the ``[max_corrs,4]`` rows each variant
returns after ``torch.manual_seed(seed)`` (the TOYL variant draws the source subsample first when N1 > 1000).
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
from utils.misc import torch_sample_select  # noqa: E402  (reference)
from utils.pcd import pdist  # noqa: E402  (reference)

from oryon_b200 import synth  # noqa: E402


def reference_kp_matcher(script: str):
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "scripts", "evaluation", script)).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def nn_correspondences"))
    end = next(i for i, l in enumerate(src) if l.startswith("@hydra.main"))
    env = {"torch": torch, "Tensor": torch.Tensor, "pdist": pdist, "torch_sample_select": torch_sample_select}
    exec("\n".join(src[start:end]), env)
    return env["nn_correspondences"]


def sift_sets(case: str):
    import cv2 as cv
    hw, seed, related, _ = synth.SIFT_CASES[case]
    img_a, img_q = synth.textured_frame_pair(seed, hw, related)
    sift = cv.SIFT_create()
    out = []
    for img in (img_a, img_q):
        kp, feats = sift.detectAndCompute(img, None)
        feats = np.asarray(feats)
        assert np.array_equal(feats, np.round(feats)) and feats.min() >= 0 and feats.max() <= 255
        out += [feats.astype(np.uint8), np.asarray([k.pt for k in kp]).reshape(-1, 2).astype(np.int16)]
    return out


def main():
    nocs_fn, toyl_fn = reference_kp_matcher("sift_nocs.py"), reference_kp_matcher("sift_toyl.py")
    for case in synth.SIFT_CASES:
        f1, k1, f2, k2 = sift_sets(case)
        t1, t2 = torch.tensor(f1.astype(np.float32)), torch.tensor(f2.astype(np.float32))
        tk1, tk2 = torch.tensor(k1), torch.tensor(k2)
        th = synth.SIFT_CASES[case][3]
        dist = torch.cat([pdist(t1[r:r + 256], t2, "inv_norm_cosine") for r in range(0, t1.shape[0], 256)])
        d64 = 0.5 * (1 - torch.nn.functional.normalize(t1.double(), dim=1) @ torch.nn.functional.normalize(t2.double(), dim=1).T)
        top2 = torch.topk(d64, k=2, dim=1, largest=False)[0]
        rec = dict(f1=f1, k1=k1, f2=f2, k2=k2, nn_idx=torch.argmin(dist, dim=1).numpy(), min_dist=torch.amin(dist, dim=1).numpy(),
                   margin=(top2[:, 1] - top2[:, 0]).float().numpy(), seed=np.int64(synth.SIFT_CASES[case][1] + 40), threshold=np.float64(th))
        n_valid = int((torch.amin(dist, dim=1) < th).sum())
        torch.manual_seed(int(rec["seed"]))
        rec["corrs_toyl"] = toyl_fn(t1, t2, tk1, tk2, th, 500).numpy()
        torch.manual_seed(int(rec["seed"]))
        try:
            rec["corrs_nocs"] = nocs_fn(t1, t2, tk1, tk2, th, 500).numpy()
            rec["nocs_raises"] = np.bool_(False)
        except RuntimeError as e:           # multinomial on an empty match set
            rec["nocs_raises"] = np.bool_(True)
            print("   nocs variant raises:", str(e)[:80])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"sift_kp_{case}.npz"), **rec)
        print(case, "N1", f1.shape[0], "N2", f2.shape[0], "rows below threshold", n_valid, "toyl rows", rec["corrs_toyl"].shape)


if __name__ == "__main__":
    main()
