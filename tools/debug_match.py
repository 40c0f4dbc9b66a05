"""Diagnostics for the matcher on a GPU box: compares both library modes against a torch fp32 evaluation
and prints where they differ.  Development aid, not part of the product or the tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oryon_b200 import _lib, synth
from oryon_b200.utils import pcd


def ref(fa, fq):
    a = torch.nn.functional.normalize(fa.flatten(2).transpose(1, 2), dim=2)
    q = torch.nn.functional.normalize(fq.flatten(2).transpose(1, 2), dim=2)
    s = a @ q.transpose(1, 2)
    v, i = s.max(2)
    return i, 0.5 * (1 - v)


torch.backends.cuda.matmul.allow_tf32 = False
for (B, D, h, w) in [(1, 32, 16, 16), (1, 32, 32, 32), (2, 32, 48, 48), (1, 64, 32, 32), (1, 128, 40, 40), (2, 256, 24, 24), (1, 17, 20, 30),
                     (4, 128, 120, 160)]:
    fa, fq, perm = synth.permuted_feature_batch(1, B, D, h, w, noise=0.3, device="cuda")
    ri, rd = ref(fa, fq)
    for mode in (_lib.MATCH_EXACT_FP32, _lib.MATCH_TC_REFINED):
        try:
            idx, dist = pcd.match_nn(fa, fq, mode=mode)
            torch.cuda.synchronize()
        except Exception as e:
            print("FAIL", (B, D, h, w), mode, e)
            raise
        bad = (idx.long() != ri)
        print(f"B{B} D{D} {h}x{w} mode{mode}: idx mismatches {int(bad.sum())}/{bad.numel()}  max|dist diff| {float((dist - rd).abs().max()):.3e}  "
              f"stats {pcd.match_last_stats()}")
        if bad.any() and mode == 1:
            b, r = [int(x[0]) for x in torch.nonzero(bad, as_tuple=True)]
            print("   first bad row", b, r, "got", int(idx[b, r]), float(dist[b, r]), "ref", int(ri[b, r]), float(rd[b, r]))
