"""Generate ``tests/golden/ckpt_remap.json``: the key rewriting of ``Oryon.init_all`` (reference net.py:99-139),
recorded by EXECUTING the reference's own loop.

TEST INFRASTRUCTURE, build container only.  ``net.py`` cannot be imported (it needs ``clip``), so the block between
``new_state_dict = dict()`` and ``inco_keys = ...`` is cut out of the source text as it lies under /root/reference,
dedented and exec'd with a synthetic ``ckpt`` whose values are integers (only the key strings matter).
"""
import json
import os
import textwrap
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE_ROOT = os.environ.get("ORYON_REFERENCE_ROOT", "/root/reference")

KEYS = [
    "sem_seg_head.predictor.transformer.layers.0.swin_block.block_1.attn.q.weight",
    "sem_seg_head.predictor.transformer.conv1.weight",
    "sem_seg_head.predictor.transformer.guidance_projection.0.0.bias",
    "sem_seg_head.predictor.transformer.decoder_guidance_projection.1.0.weight",
    "sem_seg_head.predictor.transformer.decoder1.up.weight",
    "sem_seg_head.predictor.transformer.decoder2.conv.double_conv.0.weight",
    "sem_seg_head.predictor.transformer.head.weight",
    "sem_seg_head.predictor.transformer.head.bias",
    "sem_seg_head.predictor.transformer.header.weight",
    "sem_seg_head.predictor.transformer.layers.1.fusion.decoder.weight",
    "sem_seg_head.predictor.clip_model.visual.conv1.weight",
    "sem_seg_head.predictor.clip_model.transformer.resblocks.3.attn.in_proj_weight",
    "sem_seg_head.predictor.clip_model.logit_scale",
    "sem_seg_head.predictor.upsample1.weight",
    "sem_seg_head.predictor.transformer_extra.weight",
    "sem_seg_head.pixel_decoder.weight",
    "backbone.layers.0.blocks.0.attn.qkv.weight",
    "x.sem_seg_head.predictor.transformer.conv1.weight",
]


def reference_remap(keys, vlm):
    src = open(os.path.join(REFERENCE_ROOT, "net.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if "new_state_dict = dict()" in l)
    end = next(i for i, l in enumerate(src) if "inco_keys = self.load_state_dict" in l)
    block = textwrap.dedent("\n".join(src[start:end]))
    self = types.SimpleNamespace(args=types.SimpleNamespace(image_encoder=types.SimpleNamespace(vlm=vlm)))
    env = {"ckpt": {"model": {k: i for i, k in enumerate(keys)}}, "self": self}
    exec(block, env)
    return env["new_state_dict"]


def main():
    gold = {"keys": KEYS, "clip": reference_remap(KEYS, "clip"), "other_vlm": reference_remap(KEYS, "dino")}
    with open(os.path.join(ROOT, "tests", "golden", "ckpt_remap.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print(json.dumps(gold["clip"], indent=1))


if __name__ == "__main__":
    main()
