#!/usr/bin/env python
"""Quick GPU check of the fused attention kernel against the materialised path (two batched GEMMs + softmax kernel, selected per
PROCESS by ORYON_ATTN_MATERIALIZED=1): `python tools/attn_check.py save` writes the network outputs of 2 synthetic pairs,
`ORYON_ATTN_MATERIALIZED=1 python tools/attn_check.py compare` prints the maximum absolute differences against them.  The parity
test proper is tests/test_backbone_gpu.py (against the float32 CPU oracle, minutes); this takes seconds."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oryon_b200 import synth_backbone as sb  # noqa: E402
from oryon_b200.net import Oryon  # noqa: E402

PATH = os.path.join(ROOT, "gpurun_out", "attn_check_outputs.pt")


def main():
    torch.cuda.set_device(0)
    B = 2
    model = Oryon(None, "cuda:0", state_dict=sb.oryon_state_dict(11))
    rgb_a, rgb_q = sb.synthetic_images(1, B).cuda(), sb.synthetic_images(2, B).cuda()
    emb = model.encode_tokens(sb.synthetic_tokens(3, 1)[0].cuda())[None].expand(B, -1, -1).contiguous()
    out, dbg = model.forward_tensors(rgb_a, rgb_q, emb, return_debug=True)
    torch.cuda.synchronize()
    res = {k: v.cpu() for k, v in {**out, "clip_tokens": dbg["clip_tokens"]}.items()}
    if sys.argv[1:] == ["save"]:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        torch.save(res, PATH)
        print(json.dumps({"saved": PATH, "finite": all(bool(torch.isfinite(v).all()) for v in res.values())}))
    else:
        ref = torch.load(PATH)
        print(json.dumps({"materialized": os.environ.get("ORYON_ATTN_MATERIALIZED") is not None,
                          "max_abs_diff": {k: float((res[k] - ref[k]).abs().max()) for k in ref},
                          "max_abs": {k: float(ref[k].abs().max()) for k in ref}}))


if __name__ == "__main__":
    main()
