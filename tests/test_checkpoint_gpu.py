"""GPU: N4 checkpoint ingest end to end.  A synthetic REFERENCE-LAYOUT file tree -- the OpenAI CLIP download
in ``~/.cache/clip`` (float16, as released), torchvision's ``swin_b`` file in the torch hub cache, ``pretrained_models/catseg.pth``
with CATSeg's key prefixes, a Lightning ``.ckpt`` with ``model.*`` entries, the PointDSC snapshot folder (``config.json`` +
``models/model_best.pkl``) -- is read the way the reference's constructors read it (net.py:27-34, :99-139; run_test.py:42;
utils/pointdsc/init.py:32-57): ``Oryon(args, device)`` and ``get_pointdsc_solver(path, device)``.  The loaded model must
reproduce, bit for bit, the outputs of the same network fed the expected final ``state_dict`` directly, where "expected" is
assembled here by hand in the reference's load order (stock CLIP < CATSeg's CLIP and fusion / decoder < Lightning checkpoint).
"""
import json
import os

import pytest
import torch

from gpu_util import need_gpu
from oryon_b200 import _lib, synth
from oryon_b200 import synth_backbone as sb
from oryon_b200.net import Oryon
from oryon_b200.utils.pointdsc import init as pdsc

pytestmark = pytest.mark.gpu

VIS, TXT = 2, 2      # layers of the two CLIP towers in this test (file sizes; the ingest path does not depend on the depth)


def _write_layout(home, root):
    """Returns (args, expected final state_dict)."""
    final = sb.oryon_state_dict(11, vis_layers=VIS, txt_layers=TXT)
    other = sb.oryon_state_dict(12, vis_layers=VIS, txt_layers=TXT)
    clip_keys = [k for k in final if k.startswith("vlm.clip_model.")]
    fd_keys = [k for k in final if k.startswith(("fusion.", "decoder."))]
    # 1. stock CLIP download: float16 values of ANOTHER seed (CATSeg's fine-tuned CLIP must win over it)
    stock = {k[len("vlm.clip_model."):]: other[k].to(torch.float16) for k in clip_keys}
    stock.update(input_resolution=torch.tensor(336), context_length=torch.tensor(77), vocab_size=torch.tensor(49408))
    f = os.path.join(home, ".cache", "clip", "ViT-L-14-336px.pt")
    os.makedirs(os.path.dirname(f), exist_ok=True)
    torch.save(stock, f)
    # 2. torchvision swin_b: the whole model's file; only features.0-4 feed the guidance nodes
    from torchvision.models import swin_b
    full = swin_b(weights=None).state_dict()
    for k in list(full):
        if "guidance_backbone." + k in final:
            full[k] = final["guidance_backbone." + k].clone()
    f = os.path.join(home, ".cache", "torch", "hub", "checkpoints", "swin_b-68c6b09e.pth")
    os.makedirs(os.path.dirname(f), exist_ok=True)
    torch.save(full, f)
    # 3. CATSeg checkpoint: its own prefixes; carries the final CLIP and the fusion; decoder entries of the OTHER seed
    catseg = {"sem_seg_head.predictor.clip_model." + k[len("vlm.clip_model."):]: final[k] for k in clip_keys}
    for k in fd_keys:
        if k.startswith("fusion."):
            catseg["sem_seg_head.predictor.transformer." + k[len("fusion."):]] = final[k]
        elif k.startswith("decoder.decoder"):
            catseg["sem_seg_head.predictor.transformer.decoder" + k[len("decoder.decoder"):]] = other[k]
        elif k.startswith("decoder.head"):
            catseg["sem_seg_head.predictor.transformer.head" + k[len("decoder.head"):]] = other[k]
    catseg["backbone.unrelated.weight"] = torch.zeros(3)
    os.makedirs(os.path.join(root, "pretrained_models"), exist_ok=True)
    torch.save({"model": catseg}, os.path.join(root, "pretrained_models", "catseg.pth"))
    # 4. Lightning checkpoint: the trained decoder (wins last), plus entries that are not the network's
    light = {"model." + k: final[k] for k in fd_keys if k.startswith("decoder.")}
    light.update({"feature_loss.weight": torch.zeros(1), "pointdsc_solver.sigma": torch.ones(1)})
    os.makedirs(os.path.join(root, "ckpts"), exist_ok=True)
    torch.save({"state_dict": light, "epoch": 3}, os.path.join(root, "ckpts", "last.ckpt"))
    args = dict(model=dict(use_catseg_ckpt=True, image_encoder=dict(vlm="clip", img_size=[192, 192])), eval=dict(ckpt="ckpts/last.ckpt"))
    return args, final


def test_reference_layout_files_to_loaded_model(tmp_path, monkeypatch):
    need_gpu()
    home, root = str(tmp_path / "home"), str(tmp_path / "repo")
    args, final = _write_layout(home, root)
    monkeypatch.setenv("HOME", home)
    monkeypatch.chdir(root)
    rgb_a, rgb_q = sb.synthetic_images(1, 2), sb.synthetic_images(2, 2)
    tokens = sb.synthetic_tokens(3, 1)[0]
    _lib.destroy_all()            # a library handle holds ONE packed network: a fresh handle per model of this test
    direct = Oryon(None, "cuda:0", state_dict=final, vis_layers=VIS, txt_layers=TXT)
    emb_d = direct.encode_tokens(tokens)
    out_d = {k: v.clone() for k, v in direct.forward_tensors(rgb_a, rgb_q, emb_d[None].expand(2, -1, -1).contiguous()).items()}
    emb_d = emb_d.clone()
    torch.cuda.synchronize()
    _lib.destroy_all()
    from_files = Oryon(args, "cuda:0", vis_layers=VIS, txt_layers=TXT)       # reads the four files like the reference's constructor + ckpt load
    assert from_files._load_error is None and from_files._loaded
    emb_f = from_files.encode_tokens(tokens)
    out_f = from_files.forward_tensors(rgb_a, rgb_q, emb_f[None].expand(2, -1, -1).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(emb_f, emb_d)
    for k in ("featmap_a", "featmap_q", "mask_a", "mask_q"):
        assert torch.isfinite(out_f[k]).all() and torch.equal(out_f[k], out_d[k]), k
    # a wrong merge order is visible: the stock CLIP / CATSeg's decoder alone give other outputs
    wrong = dict(final)
    other = sb.oryon_state_dict(12, vis_layers=VIS, txt_layers=TXT)
    wrong.update({k: other[k] for k in final if k.startswith("decoder.")})
    out_f = {k: v.clone() for k, v in out_f.items()}
    _lib.destroy_all()
    out_w = Oryon(None, "cuda:0", state_dict=wrong, vis_layers=VIS, txt_layers=TXT).forward_tensors(rgb_a, rgb_q, emb_d[None].expand(2, -1, -1).contiguous())
    assert not torch.equal(out_w["featmap_a"], out_d["featmap_a"])
    torch.cuda.synchronize()
    _lib.destroy_all()


def test_missing_file_is_named_by_the_first_forward(tmp_path, monkeypatch):
    need_gpu()
    home, root = str(tmp_path / "home"), str(tmp_path / "repo")
    args, _ = _write_layout(home, root)
    os.remove(os.path.join(root, "pretrained_models", "catseg.pth"))
    monkeypatch.setenv("HOME", home)
    monkeypatch.chdir(root)
    _lib.destroy_all()
    model = Oryon(args, "cuda:0", vis_layers=VIS, txt_layers=TXT)
    with pytest.raises(RuntimeError, match="catseg.pth"):
        model.forward_tensors(sb.synthetic_images(1, 1), sb.synthetic_images(2, 1), torch.zeros(1, 80, 768))


def test_pointdsc_snapshot_folder_to_solver(tmp_path):
    """``get_pointdsc_solver(ckpt_path, device)`` on the released folder layout (utils/pointdsc/init.py:32-57): hyper-parameters
    from ``config.json`` (``nms_radius`` <- ``inlier_threshold``, :49), weights from ``models/model_best.pkl``."""
    need_gpu()
    cfg = dict(synth.POINTDSC_DEFAULT_CFG)
    sd = synth.pointdsc_state_dict(300)
    snap = tmp_path / "pointdsc" / "snapshot" / "PointDSC_3DMatch_release"
    (snap / "models").mkdir(parents=True)
    json.dump({**cfg, "dataset": "3DMatch", "unused_training_key": 1}, open(snap / "config.json", "w"))
    torch.save(sd, snap / "models" / "model_best.pkl")
    solver = pdsc.get_pointdsc_solver(str(tmp_path / "pointdsc"), "cuda:0")
    assert solver.num_layers == cfg["num_layers"] and solver.nms_radius == cfg["inlier_threshold"] and solver.k == cfg["k"]
    direct = pdsc.PointDSCSolver(sd, in_dim=cfg["in_dim"], num_layers=cfg["num_layers"], num_channels=cfg["num_channels"],
                                 num_iterations=cfg["num_iterations"], ratio=cfg["ratio"], sigma_d=cfg["sigma_d"], k=cfg["k"],
                                 nms_radius=cfg["inlier_threshold"], device="cuda:0")
    data = synth.rigid_correspondences(301, n=500, outlier_frac=0.3)
    T_files = pdsc.get_pointdsc_pose(solver, data["src"], data["tgt"], "cuda:0")
    T_direct = pdsc.get_pointdsc_pose(direct, data["src"], data["tgt"], "cuda:0")
    assert torch.equal(T_files, T_direct) and T_files.shape == (4, 4) and bool(torch.isfinite(T_files).all())
    R = T_files[:3, :3].double()
    assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-4)      # a rigid motion came out
