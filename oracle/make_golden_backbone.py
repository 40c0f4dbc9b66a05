"""Golden fixtures for the fusion + decoder restatement: runs the UNMODIFIED reference modules
(models/fusion.py ImageTextFusion, models/decoder.py StandardDecoder) from /root/reference on seeded inputs
and weights and commits small samples of their outputs under tests/golden/backbone_fd_<seed>.npz.

TEST INFRASTRUCTURE; build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_backbone.py
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

from models.decoder import StandardDecoder  # noqa: E402  (reference)
from models.fusion import ImageTextFusion  # noqa: E402  (reference)

from oryon_b200 import synth, synth_backbone  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SAMPLE_STRIDE = 997


def fd_inputs(seed: int, b: int = 2):
    g = torch.Generator().manual_seed(seed)
    img_feats = torch.randn(b, 1024, 24, 24, generator=g)
    text = torch.randn(b, 1, 80, 768, generator=g)
    guid = [torch.randn(b, 512, 24, 24, generator=g), torch.randn(b, 256, 48, 48, generator=g), torch.randn(b, 128, 96, 96, generator=g)]
    return img_feats, text, guid


def main():
    torch.set_num_threads(8)
    for seed in (700, 701):
        sd = synth_backbone.fusion_decoder_state_dict(seed)
        fusion = ImageTextFusion("cpu")
        decoder = StandardDecoder("cpu", True, True, input_dim=128, decoder_dims=[64, 32])
        fsd = {k[len("fusion."):]: v for k, v in sd.items() if k.startswith("fusion.")}
        for k, v in fusion.state_dict().items():  # attn_mask buffers are derived, not parameters
            if k.endswith("attn_mask"):
                fsd[k] = v
        fusion.load_state_dict(fsd, strict=True)
        decoder.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
        fusion.eval(), decoder.eval()
        img_feats, text, guid = fd_inputs(seed)
        with torch.no_grad():
            f = fusion.forward(img_feats, text, guid)
            logits, featmap = decoder.forward(f, guid)
        np.savez_compressed(os.path.join(OUT, f"backbone_fd_{seed}.npz"),
                            in_sum=np.array([synth.tensor_checksum(img_feats), synth.tensor_checksum(guid[2]),
                                             synth.tensor_checksum(sd["fusion.conv1.weight"])]),
                            fusion=f.flatten()[::SAMPLE_STRIDE].numpy(), logits=logits.flatten()[::SAMPLE_STRIDE].numpy(),
                            featmap=featmap.flatten()[::SAMPLE_STRIDE].numpy(),
                            stats=np.array([f.mean(), f.std(), logits.mean(), logits.std(), featmap.mean(), featmap.std(),
                                            (logits > 0).float().mean()]))
        print(seed, "fusion", tuple(f.shape), float(f.std()), "logits", tuple(logits.shape), float(logits.std()),
              "featmap", tuple(featmap.shape), float(featmap.std()), "mask frac", float((logits > 0).float().mean()))


if __name__ == "__main__":
    main()
