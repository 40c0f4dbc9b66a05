#!/usr/bin/env python
"""Which GEMMs of the network need three tcgen05 products?  The network runs once with every GEMM at
precision 3 (fp16 split pairs: hi*hi + lo*hi + hi*lo, float32-equivalent), then once per GEMM class with THAT class demoted to
a single fp16 product (ORYON_GEMM_P1_SHAPES, gemm.cu) and everything else unchanged; the maximum absolute change of the
outputs (feature maps, mask logits) is what demoting the class would add to the distance from the float32 reference, to be held
against the 1e-3 gate (every stage sits at <= 3e-4 with three products everywhere, tests/test_backbone_gpu.py).

    gpurun -- 'python tools/gemm_precision_sweep.py > gpurun_out/r02_gemm_precision_sweep.json'
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oryon_b200 import synth_backbone as sb  # noqa: E402
from oryon_b200.net import Oryon  # noqa: E402

# (N, K) classes of one network pass; see tools/bench_gemm_all.py for the shapes
CLASSES = {
    "clip_qkv (24 layers)": "3072:1024", "clip_out_proj (24)": "1024:1024", "clip_mlp_fc (24)": "4096:1024", "clip_mlp_proj (24)": "1024:4096",
    "clip_patch_embed": "1024:588",
    "all CLIP linear layers": "3072:1024,1024:1024,4096:1024,1024:4096,1024:588",
    "everything EXCEPT the CLIP linear layers (swin, fusion, decoder, convs)": "!3072:1024,1024:1024,4096:1024,1024:4096,1024:588",
    "swin stage 1 (qkv, proj, mlp)": "384:128,128:128,512:128,128:512,128:48",
    "swin stage 2 + merges": "768:256,256:256,1024:256,256:1024,256:512,512:1024",
    "fusion (clip_conv, cost volume, 7x7 conv, guidance, swin blocks)": "768:1024,80:768,128:3920,128:4608,384:256,128:128,512:128,128:512",
    "decoder (ConvT as GEMM + 3x3 convs)": "32:2304,16:1152,384:128,64:1152,64:576,192:64,32:576,32:288,128:32",
}


def main():
    torch.cuda.set_device(0)
    B = 2
    model = Oryon(None, "cuda:0", state_dict=sb.oryon_state_dict(11))
    rgb_a, rgb_q = sb.synthetic_images(1, B).cuda(), sb.synthetic_images(2, B).cuda()
    emb = model.encode_tokens(sb.synthetic_tokens(3, 1)[0].cuda())[None].expand(B, -1, -1).contiguous()

    def run():
        out = model.forward_tensors(rgb_a, rgb_q, emb)
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in out.items()}

    os.environ.pop("ORYON_GEMM_P1_SHAPES", None)
    ref = run()
    again = run()
    res = {"gate": 1e-3, "note": "max |output(class at 1 product) - output(all at 3 products)| over 2 pairs; seeded random weights",
           "rerun_noise": {k: float((again[k] - ref[k]).abs().max()) for k in ref}, "classes": {}}
    for name, spec in CLASSES.items():
        os.environ["ORYON_GEMM_P1_SHAPES"] = spec
        out = run()
        res["classes"][name] = {"shapes": spec, **{k: float((out[k] - ref[k]).abs().max()) for k in ref}}
    os.environ.pop("ORYON_GEMM_P1_SHAPES", None)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
