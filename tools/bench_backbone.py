"""Per-kernel timing of the network + post-network path on synthetic full-size inputs (not the contract bench)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oryon_b200 import _lib, synth_backbone as sb  # noqa: E402
from oryon_b200.net import Oryon, gemm_counters  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--precision", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--text", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    t0 = time.time()
    w = sb.oryon_state_dict(11)
    model = Oryon(None, "cuda:0", state_dict=w, precision=a.precision, max_pairs_per_pass=a.chunk)
    del w
    print(f"weights generated + loaded in {time.time() - t0:.1f}s", flush=True)
    B = a.pairs
    rgb_a, rgb_q = sb.synthetic_images(1, B).cuda(), sb.synthetic_images(2, B).cuda()
    tokens = sb.synthetic_tokens(3, 1).cuda()
    emb1 = model.encode_tokens(tokens[0])
    emb = emb1[None].expand(B, -1, -1).contiguous()
    out = {}
    if a.text:
        for _ in range(2):
            model.encode_tokens(tokens[0])
        torch.cuda.synchronize()
        _lib.profile_enable(0, True)
        gemm_counters(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            model.encode_tokens(tokens[0])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        prof = _lib.profile_read(0)
        n, fl = gemm_counters(0)
        out["text_80_prompts_ms"] = ms
        out["text_kernels_ms"] = {str(k): v[0] / a.steps for k, v in prof.items()}
        out["text_gemm_tflops"] = fl / a.steps / (prof["gemm_tc"][0] / a.steps * 1e-3) / 1e12
        _lib.profile_enable(0, False)
    for _ in range(2):
        model.forward_tensors(rgb_a, rgb_q, emb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model.forward_tensors(rgb_a, rgb_q, emb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    _lib.profile_enable(0, True)
    gemm_counters(0)
    for _ in range(a.steps):
        model.forward_tensors(rgb_a, rgb_q, emb)
    torch.cuda.synchronize()
    prof = _lib.profile_read(0)
    n, fl = gemm_counters(0)
    _lib.profile_enable(0, False)
    out.update(pairs=B, precision=a.precision, network_ms=ms, pairs_per_s=B / ms * 1e3,
               kernels_ms={str(k): v[0] / a.steps for k, v in prof.items()}, kernel_launches={str(k): v[1] / a.steps for k, v in prof.items()},
               gemm_launches=n / a.steps, gemm_algorithmic_tflop=fl / a.steps / 1e12,
               gemm_tflops=fl / a.steps / (prof["gemm_tc"][0] / a.steps * 1e-3) / 1e12,
               workspace_gb=_lib.load().oryon_workspace_bytes(_lib.handle(0)) / 1e9)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
