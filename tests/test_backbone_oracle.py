"""CPU: the network oracle (oracle/backbone_oracle.py) is pinned
  * fusion + decoder: against outputs of the UNMODIFIED reference modules (tests/golden/backbone_fd_*.npz,
    written by oracle/make_golden_backbone.py from models/fusion.py and models/decoder.py);
  * CLIP towers (third-party clip==1.0, absent): against the independent implementation in `transformers`
    (CLIPVisionModel / CLIPTextModel) with the same weights, small configuration;
  * guidance backbone: it IS torchvision's swin_b; checked for the reference's output shapes (net.py:50-54).
"""
import os

import numpy as np
import pytest
import torch

import backbone_oracle as bo
from oryon_b200 import synth, synth_backbone

STRIDE = 997


def _fd_inputs(seed, b=2):
    g = torch.Generator().manual_seed(seed)
    img_feats = torch.randn(b, 1024, 24, 24, generator=g)
    text = torch.randn(b, 1, 80, 768, generator=g)
    guid = [torch.randn(b, 512, 24, 24, generator=g), torch.randn(b, 256, 48, 48, generator=g), torch.randn(b, 128, 96, 96, generator=g)]
    return img_feats, text, guid


@pytest.mark.parametrize("seed", [700, 701])
def test_fusion_decoder_oracle_equals_reference(golden_dir, seed):
    g = np.load(os.path.join(golden_dir, f"backbone_fd_{seed}.npz"))
    sd = synth_backbone.fusion_decoder_state_dict(seed)
    img_feats, text, guid = _fd_inputs(seed)
    assert [synth.tensor_checksum(t) for t in (img_feats, guid[2], sd["fusion.conv1.weight"])] == list(g["in_sum"])
    torch.set_num_threads(8)
    with torch.no_grad():
        f = bo.fusion_forward(sd, img_feats, text, guid[0])
        logits, featmap = bo.decoder_forward(sd, f, guid)
    assert f.shape == (2, 128, 1, 24, 24) and logits.shape == (2, 1, 192, 192) and featmap.shape == (2, 32, 192, 192)
    np.testing.assert_allclose(f.flatten()[::STRIDE].numpy(), g["fusion"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(logits.flatten()[::STRIDE].numpy(), g["logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(featmap.flatten()[::STRIDE].numpy(), g["featmap"], rtol=1e-4, atol=2e-5)


def _hf_to_openai_blocks(hf_layers, prefix, out):
    for i, l in enumerate(hf_layers):
        p = f"{prefix}.resblocks.{i}"
        a = l.self_attn
        out[p + ".attn.in_proj_weight"] = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]).detach()
        out[p + ".attn.in_proj_bias"] = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]).detach()
        out[p + ".attn.out_proj.weight"], out[p + ".attn.out_proj.bias"] = a.out_proj.weight.detach(), a.out_proj.bias.detach()
        out[p + ".ln_1.weight"], out[p + ".ln_1.bias"] = l.layer_norm1.weight.detach(), l.layer_norm1.bias.detach()
        out[p + ".ln_2.weight"], out[p + ".ln_2.bias"] = l.layer_norm2.weight.detach(), l.layer_norm2.bias.detach()
        out[p + ".mlp.c_fc.weight"], out[p + ".mlp.c_fc.bias"] = l.mlp.fc1.weight.detach(), l.mlp.fc1.bias.detach()
        out[p + ".mlp.c_proj.weight"], out[p + ".mlp.c_proj.bias"] = l.mlp.fc2.weight.detach(), l.mlp.fc2.bias.detach()


def test_clip_restatement_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModel
    torch.manual_seed(0)
    vcfg = CLIPVisionConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4, image_size=56,
                            patch_size=14, hidden_act="quick_gelu", attn_implementation="eager")
    vm = CLIPVisionModel(vcfg).eval()
    for prm in vm.parameters():  # default init is tiny: make every term matter
        prm.data = torch.randn_like(prm) * (0.2 if prm.dim() > 1 else 0.5) + (1.0 if "norm" in "" else 0.0)
    w = {}
    e = vm.vision_model.embeddings
    pv = "vlm.clip_model.visual"
    w[pv + ".conv1.weight"] = e.patch_embedding.weight.detach()
    w[pv + ".class_embedding"] = e.class_embedding.detach()
    w[pv + ".positional_embedding"] = e.position_embedding.weight.detach()
    w[pv + ".ln_pre.weight"], w[pv + ".ln_pre.bias"] = vm.vision_model.pre_layrnorm.weight.detach(), vm.vision_model.pre_layrnorm.bias.detach()
    w[pv + ".ln_post.weight"], w[pv + ".ln_post.bias"] = torch.ones(64), torch.zeros(64)
    _hf_to_openai_blocks(vm.vision_model.encoder.layers, pv + ".transformer", w)
    img = torch.rand(2, 3, 37, 37)
    with torch.no_grad():
        pre = bo.clip_preprocess(img, 56)
        hf = vm(pixel_values=pre).last_hidden_state[:, 1:, :]           # tokens before post_layernorm
        hf = torch.nn.functional.layer_norm(hf, (64,))
        mine = bo.clip_encode_image(w, img, layers=3, heads=4, patch=14, size=56)
    np.testing.assert_allclose(mine.flatten(2).transpose(1, 2).numpy(), hf.numpy(), rtol=1e-4, atol=1e-4)

    tcfg = CLIPTextConfig(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4,
                          max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=999, attn_implementation="eager")
    tm = CLIPTextModel(tcfg).eval()
    for prm in tm.parameters():
        prm.data = torch.randn_like(prm) * (0.2 if prm.dim() > 1 else 0.5)
    pt = "vlm.clip_model"
    w[pt + ".token_embedding.weight"] = tm.text_model.embeddings.token_embedding.weight.detach()
    w[pt + ".positional_embedding"] = tm.text_model.embeddings.position_embedding.weight.detach()
    w[pt + ".ln_final.weight"], w[pt + ".ln_final.bias"] = tm.text_model.final_layer_norm.weight.detach(), tm.text_model.final_layer_norm.bias.detach()
    w[pt + ".text_projection"] = torch.eye(64)
    _hf_to_openai_blocks(tm.text_model.encoder.layers, pt + ".transformer", w)
    toks = synth_backbone.synthetic_tokens(3, 1, 5, 77, 1000)[0]
    with torch.no_grad():
        hf = tm(input_ids=toks).pooler_output                            # final LN at the EOT (highest id) position
        mine = bo.clip_encode_tokens(w, toks, layers=3, heads=4)
    np.testing.assert_allclose(mine.numpy(), hf.numpy(), rtol=1e-4, atol=1e-4)


def test_guidance_backbone_shapes():
    sd = synth_backbone.swin_state_dict(5)
    net = bo.guidance_backbone(sd)
    g = bo.guidance_embeds(net, synth_backbone.synthetic_images(1, 1))
    assert [tuple(t.shape) for t in g] == [(1, 512, 24, 24), (1, 256, 48, 48), (1, 128, 96, 96)]
    assert all(torch.isfinite(t).all() and 0.05 < t.std() < 50 for t in g)
