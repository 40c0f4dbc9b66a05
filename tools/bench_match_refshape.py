#!/usr/bin/env python
"""The matcher where the reference lives: D = 32, 192 x 192 feature maps, N1 <= 5000 anchor rows against
N2 up to 36 864 query positions, B = 32 pairs per launch, on SMOOTH maps -- where neighbouring columns are near-ties and the
candidate lists of the tensor-core pass get long -- next to the iid best case.  Prints one JSON document: per case the step
time, the per-kernel split (library CUDA events), the candidate-list histograms (oryon_match_list_hist), overflow rows.

    gpurun -- 'python tools/bench_match_refshape.py > gpurun_out/r02_match_refshape.json'

Cases: `smooth_c8 / c4 / c16` = synth.smooth_feature_pair with coarse grids of 8 / 4 / 16 feature-map pixels (bilinear, +2 %
white noise; the query is the rolled anchor + 5 % noise); `network` = the decoder outputs of the seeded random-weight network
on synthetic pairs (ConvT + bilinear smooth, the closest thing to real maps available offline); `iid` = white-noise features
with a planted permutation (config 2's distribution at this shape).  Query side: an ellipse of ~5300 px (`roi`) or the whole
map (`full`, 36 864)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oryon_b200 import _lib, synth  # noqa: E402
from oryon_b200.utils import pcd  # noqa: E402


def run_case(name, fa, fq, ma, mq, reps=20):
    dev = fa.device
    B = fa.shape[0]
    roi_a, cnt_a = pcd.mask_to_roi(ma)
    roi_q, cnt_q = pcd.mask_to_roi(mq)
    n_a, n_q = cnt_a.tolist(), cnt_q.tolist()
    for _ in range(3):
        idx, dist = pcd.match_nn(fa, fq, roi_a, roi_q, n_a, n_q)
    torch.cuda.synchronize()
    _lib.profile_enable(dev.index, True)
    _lib.profile_read(dev.index)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx, dist = pcd.match_nn(fa, fq, roi_a, roi_q, n_a, n_q)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read(dev.index)
    _lib.profile_enable(dev.index, False)
    stats = pcd.match_last_stats(dev)
    pcd.match_set_hist(True, dev)
    pcd.match_nn(fa, fq, roi_a, roi_q, n_a, n_q)
    hist = pcd.match_list_hist(dev)
    pcd.match_set_hist(False, dev)
    # exactness on a sample of rows of pair 0 against float64 (the reference computes float32: ties to 1e-6 are distance-equivalent)
    b = 0
    A = torch.nn.functional.normalize(fa[b].flatten(1)[:, roi_a[b, :n_a[b]].long()].T.double(), dim=1)
    Q = torch.nn.functional.normalize(fq[b].flatten(1)[:, roi_q[b, :n_q[b]].long()].T.double(), dim=1)
    rows = torch.arange(0, n_a[b], max(1, n_a[b] // 512), device=dev)
    S = A[rows] @ Q.T
    best = S.max(1).values
    got = S[torch.arange(rows.numel(), device=dev), idx[b, rows].long()]
    flops = sum(2.0 * a * q * fa.shape[1] for a, q in zip(n_a, n_q))
    ms = e0.elapsed_time(e1) / reps
    tc = prof.get("match_tc", (0.0, 1))
    return {"case": name, "B": B, "D": fa.shape[1], "n_a_mean": sum(n_a) / B, "n_q_mean": sum(n_q) / B, "ms_per_batch": ms,
            "pairs_per_s": B / (ms * 1e-3), "kernels_ms": {str(k): round(v[0] / reps, 4) for k, v in prof.items()},
            "match_tc_tflops": flops / (tc[0] / max(tc[1], 1) * 1e-3) / 1e12 if tc[0] else None,
            "stats": stats, "hist": hist, "max_score_gap_vs_float64": float((best - got).max())}


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, D, H, W = 32, 32, 192, 192
    ma = torch.stack([synth.ellipse_mask(H, W, 0.136, 0.5, 0.5) for _ in range(B)]).to(dev)           # ~5000 px
    mq_roi = torch.stack([synth.ellipse_mask(H, W, 0.145, 0.52, 0.47) for _ in range(B)]).to(dev)      # ~5300 px
    mq_full = torch.ones(B, H, W, dtype=torch.int32, device=dev)
    out = {"shape": f"B={B}, D={D}, {H}x{W}; anchor ROI ~5000 px", "cases": []}
    for coarse in (8, 4, 16):
        pairs = [synth.smooth_feature_pair(500 + i, D, H, W, coarse=coarse) for i in range(B)]
        fa = torch.stack([p[0] for p in pairs]).to(dev)
        fq = torch.stack([p[1] for p in pairs]).to(dev)
        out["cases"].append(run_case(f"smooth_c{coarse}_roi", fa, fq, ma, mq_roi))
        out["cases"].append(run_case(f"smooth_c{coarse}_full", fa, fq, ma, mq_full))
    fa, fq, _ = synth.permuted_feature_batch(7, B, D, H, W, noise=0.1, device="cuda:0")
    out["cases"].append(run_case("iid_roi", fa, fq, ma, mq_roi))
    out["cases"].append(run_case("iid_full", fa, fq, ma, mq_full))
    if "--no-network" not in sys.argv:
        sys.path.insert(0, ROOT)
        import bench
        pipe, model = bench.build_full_path(0, 3, "predicted")
        prompts = bench.prompt_lists()
        hb = bench.host_batches(1, B)
        t0 = time.perf_counter()
        o = pipe.forward(bench.step_batch(hb, prompts, 0, 0, B))
        torch.cuda.synchronize()
        res = pipe.mask_results(hb[0], o)
        out["cases"].append(run_case("network_predicted_masks", o["featmap_a"], o["featmap_q"], res["mask_a"], res["mask_q"]))
        out["cases"].append(run_case("network_anchor5000_full_query", o["featmap_a"], o["featmap_q"], ma, mq_full))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
