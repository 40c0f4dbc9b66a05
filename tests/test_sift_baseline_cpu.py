"""The key-point baseline script (scripts/evaluation/sift_baseline.py = the reference's sift_nocs.py / sift_toyl.py) on a
synthetic TOYL tree whose frames are views of one texture.  On the CPU the three hot-path calls are answered by the oracles
(injected path object) and the evaluator by the oracle backend: this covers the host logic -- SIFT on the host, mask filtering,
(x, y) key-point conventions, lifting in millimetres -> metres, pose composition, failure rows, the output file.  The ``gpu``
test in tests/test_zz_sift_kp_gpu.py runs the same loop on liboryon_b200."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oryon_oracle as oracle
from oryon_b200 import synth
from oryon_b200.datasets import TOYLDataset
from oryon_b200.utils.evaluator import format_sym_set

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("oryon_sift_baseline", os.path.join(ROOT, "scripts", "evaluation", "sift_baseline.py"))
baseline = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(baseline)

HW = (240, 320)


def write_textured_toyl_tree(d: str, seed: int = 0):
    """``synth.write_toyl_tree`` at 240 x 320 with object rectangles of 45 x 55 pixels; the colour frames replaced by views of one
    synthetic texture, so that SIFT finds matching key points inside the object masks of any two frames."""
    from PIL import Image
    info = synth.write_toyl_tree(d, seed, hw=HW, mask_scale=5)
    frames = synth.textured_frames(31 + seed, HW, 6)
    n = 0
    for s in (1, 2):
        for im in range(3):
            g = frames[n].astype(np.float32)
            rgb = np.stack([g, np.clip(g * 0.9 + 10, 0, 255), np.clip(g * 1.05, 0, 255)], -1).astype(np.uint8)
            Image.fromarray(rgb).save(os.path.join(info["base"], "split", "test", f"{s:06d}", "rgb", f"{im:06d}.png"))
            n += 1
    return info


def toyl_dataset(d, info):
    args = dict(device="cuda:0", dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj="all")),
                test=dict(mask="oracle", add_description="yes"))
    return TOYLDataset(args, eval=True)


class OraclePath:
    """The three hot-path calls answered by the CPU oracle (test infrastructure)."""

    def __init__(self, seed=300):
        self.w, self.cfg = synth.pointdsc_state_dict(seed), dict(synth.POINTDSC_DEFAULT_CFG)
        self.calls = []

    def match(self, fa, fq, ka, kq, th, n, **variant):
        self.calls.append((fa.shape[0], fq.shape[0]))
        return oracle.nn_correspondences_kp(fa, fq, ka, kq, th, n, **variant)

    def lift(self, depth, K, xy):
        return oracle.lift_pcd(depth, K, xy)

    def pose(self, pa, pq):
        return oracle.pointdsc_pose(self.w, self.cfg, pa, pq).to(torch.float32)


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("toyl_sift"))
    return d, write_textured_toyl_tree(d)


def test_baseline_loop_with_oracle_path(tree, tmp_path):
    from test_evaluator_cpu import _OracleBackend
    d, info = tree
    ds = toyl_dataset(d, info)
    models, _, symms = ds.get_object_info()
    path = OraclePath()
    out = tmp_path / "sift_toyl_oracle.txt"
    torch.manual_seed(5)
    ev = baseline.run_baseline(ds, path, "toyl", "oracle", None, False, str(out),
                               pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}))
    lines = out.read_text().splitlines()
    assert len(lines) == len(ds) == len(path.calls) and len(ev.metrics["instance_id"]) == len(ds)
    assert all(a >= 5 and q >= 5 for a, q in path.calls), path.calls            # key points inside the 45 x 55 object masks
    for line, inst in zip(lines, ds.instances):
        id_a, id_q, pose = line.split(",")
        assert id_a == f"{inst[1]} {inst[2]} {inst[-1]}" and id_q == f"{inst[3]} {inst[4]} {inst[-1]}"
        T = np.asarray([float(v) for v in pose.split(" ")]).reshape(3, 4)
        assert np.isfinite(T).all() and abs(np.linalg.det(T[:, :3]) - 1.0) < 1e-3       # a rigid transform from PointDSC
    assert sum(ev.counts["Missing segm"]) == 0
    # the three-token lines are what the offline scorer's reader accepts
    from oryon_b200.utils.evaluator import dict_from_preds
    preds, _, _, iou_present = dict_from_preds(str(out))
    assert not iou_present and set(preds) == set(ev.metrics["instance_id"])


def test_baseline_registers_failures_for_empty_masks(tree, tmp_path):
    """``--mask-dir`` (the reference's masks == 'ours'): an all-zero mask file leaves no key point -> failure row, no line."""
    from PIL import Image
    from test_evaluator_cpu import _OracleBackend
    d, info = tree
    ds = toyl_dataset(d, info)
    mdir = tmp_path / "masks"
    mdir.mkdir()
    for i in range(len(ds)):
        inst = ds.instances[i]
        for scene, img in ((inst[1], inst[2]), (inst[3], inst[4])):
            item = ds.get_item(scene, img, inst[-1], "oracle")
            m = (item["mask"] == item["metadata"]["mask_ids"][0]).astype(np.uint8)
            if i == 2 and (scene, img) == (inst[3], inst[4]):      # frame (2, 1) with object 12: used by this pair only
                m[:] = 0
            Image.fromarray(m).save(str(mdir / f"{item['instance_id']}.png"))
    models, _, symms = ds.get_object_info()
    out = tmp_path / "out.txt"
    torch.manual_seed(5)
    ev = baseline.run_baseline(ds, OraclePath(), "toyl", "ours", str(mdir), False, str(out),
                               pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}))
    assert len(out.read_text().splitlines()) == len(ds) - 1 and sum(ev.counts["Missing segm"]) == 1
    assert len(ev.metrics["instance_id"]) == len(ds)


def test_baseline_loop_nocs_variant(tmp_path):
    """The NOCS form of the loop (sift_nocs.py): pairs keyed by object NAME, no source subsample, OBJ models; the frame with a
    single-pixel object has no key point inside its mask -> failure rows for the pairs that use it."""
    from PIL import Image
    from oryon_b200.datasets import NOCSDataset
    from test_evaluator_cpu import _OracleBackend
    d = str(tmp_path)
    info = synth.write_nocs_tree(d, 0, hw=HW, mask_scale=5)
    frames = synth.textured_frames(77, HW, 6)
    for n, (s, im, _) in enumerate(info["frames"]):
        g = frames[n].astype(np.float32)
        rgb = np.stack([g, np.clip(g * 0.9 + 10, 0, 255), np.clip(g * 1.05, 0, 255)], -1).astype(np.uint8)
        Image.fromarray(rgb).save(os.path.join(info["base"], f"split/real_test/scene_{s}/{im:04d}_color.png"))
    args = dict(device="cuda:0", dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj="all")),
                test=dict(mask="oracle", add_description="yes"))
    ds = NOCSDataset(args, eval=True)
    models, _, symms = ds.get_object_info()
    out = tmp_path / "sift_nocs_oracle.txt"
    torch.manual_seed(5)
    ev = baseline.run_baseline(ds, OraclePath(), "nocs", "oracle", None, False, str(out),
                               pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}))
    lines = out.read_text().splitlines()
    single_pixel = [i for i, (sa, ia, sq, iq, obj) in enumerate(info["pairs"]) if (sq, iq) == (2, 2) and obj == list(synth.NOCS_TREE_OBJECTS)[0]]
    assert len(ev.metrics["instance_id"]) == len(ds) and len(lines) == len(ds) - len(single_pixel) and len(single_pixel) >= 1
    assert sum(ev.counts["Missing segm"]) == len(single_pixel)
    assert ev.metrics["cls_id"] == [p[-1] for p in info["pairs"]]                  # object names
    assert all(l.split(",")[0].split(" ")[2] in synth.NOCS_TREE_OBJECTS for l in lines)
