"""CPU: the TOYL test-split reader (oryon_b200/utils/data/toyl.py incl. its PLY parser, oryon_b200/datasets.py:TOYLDataset)
against what the reference's own ``TOYLDataset`` / ``utils.data.toyl`` return for the same synthetic dataset tree
(``oracle/make_golden_toyl.py`` -> ``tests/golden/toyl_tree_0.*``): pair list, ids, prompts, relative poses, frames, per-frame
annotations, validity flags, object models (ASCII and binary PLY) and the BOP symmetry sets."""
import hashlib
import json
import os

import numpy as np
import pytest

from oryon_b200 import synth
from oryon_b200.datasets import TOYLDataset
from oryon_b200.utils.data import toyl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("toyl"))
    return d, synth.write_toyl_tree(d, 0)


def _dataset(tree, obj_split, mask, add_description):
    d, info = tree
    args = dict(device="cuda:0", dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj=obj_split)),
                test=dict(mask=mask, add_description=add_description))
    return TOYLDataset(args, eval=True)


def _resized(mask01, size=(224, 224)):
    H, W = mask01.shape
    ys = np.minimum(np.floor(np.arange(size[0], dtype=np.float32) * np.float32(H / size[0])).astype(np.int64), H - 1)
    xs = np.minimum(np.floor(np.arange(size[1], dtype=np.float32) * np.float32(W / size[1])).astype(np.int64), W - 1)
    return mask01[np.ix_(ys, xs)].astype(np.uint8)


@pytest.mark.parametrize("obj_split,mask,desc", [("all", "predicted", "yes"), ("cars", "oracle", "wrong")])
def test_samples_match_the_reference_reader(tree, obj_split, mask, desc):
    gold = json.load(open(os.path.join(GOLD, "toyl_tree_0.json")))[obj_split]
    ds = _dataset(tree, obj_split, mask, desc)
    assert len(ds) == gold["length"] and ds.tracked_instances == gold["tracked"]
    for i, want in enumerate(gold["samples"]):
        item_a, item_q, prompt, pose, obj_id, instance_id, valid = ds[i]
        assert (instance_id, obj_id, bool(valid)) == (want["instance_id"], want["obj_id"], want["valid"])
        assert prompt == want["prompt"]
        np.testing.assert_array_equal(pose, np.asarray(want["pose"]))          # translation /1000: same float64 operation
        for item, w in ((item_a, want["anchor"]), (item_q, want["query"])):
            md = item["metadata"]
            assert item["instance_id"] == w["instance_id"] and list(item["mask"].shape) == w["hw_size"]
            assert (md["mask_ids"], md["cls_ids"], md["cls_names"], md["cls_descs"]) == (w["mask_ids"], w["cls_ids"], w["cls_names"], w["cls_descs"])
            assert len(md["poses"]) == w["n_poses"]
            np.testing.assert_array_equal(md["poses"][0], np.asarray(w["pose0"]))
            np.testing.assert_array_equal(item["camera"], np.asarray(w["camera"]))
            assert sha(item["rgb"]) == w["rgb_sha"] and sha(np.asarray(item["depth"]).astype(np.int64)) == w["depth_sha"]
            m224 = _resized(np.asarray(item["mask"]) == md["mask_ids"][0])
            assert sha(m224) == w["mask224_sha"] and int(m224.sum()) == w["mask224_sum"]


def test_object_models_and_symmetries_match_the_reference(tree):
    gold = json.load(open(os.path.join(GOLD, "toyl_tree_0.json")))["objects"]
    arrs = np.load(os.path.join(GOLD, "toyl_tree_0.npz"))
    models, diams, symms = _dataset(tree, "all", "oracle", "yes").get_object_info()
    assert sorted(str(k) for k in models) == sorted(gold)
    for ks, g in gold.items():
        k = int(ks)
        assert diams[k] == g["diameter"] and len(symms[k]) == g["n_symmetries"]
        for key in ("pts", "normals", "faces"):
            np.testing.assert_array_equal(models[k][key], arrs[f"{ks}/{key}"])
        np.testing.assert_array_equal(np.stack([s["R"] for s in symms[k]]), arrs[f"{ks}/sym_R"])
        np.testing.assert_array_equal(np.stack([s["t"] for s in symms[k]]), arrs[f"{ks}/sym_t"])


def test_camera_quirk_and_ids(tree):
    """``toyl.get_camera()`` returns the NOCS intrinsics in the reference too; the dataset class carries the real ones."""
    ds = _dataset(tree, "all", "predicted", "yes")
    assert toyl.get_camera()[0, 0] == 591.0125 and ds.K[0, 0] == 572.4114
    assert ds.frame_ids(1) == ("1 1 5", "2 2 5") and ds.get_obj_info("5")[1] == ds.get_object_info()[1][5]


def test_ply_parser_reads_both_encodings(tree):
    d, info = tree
    a = toyl.read_ply(os.path.join(info["base"], "models_bop", "obj_000001.ply"))      # ASCII
    b = toyl.read_ply(os.path.join(info["base"], "models_bop", "obj_000005.ply"))      # binary little endian
    for ply, nv in ((a, 10), (b, 13)):
        assert ply["vertex"]["x"].dtype == np.float32 and ply["vertex"]["x"].shape == (nv,)
        assert ply["face"]["vertex_indices"].shape == (nv - 2, 3) and ply["face"]["vertex_indices"].dtype == np.int32
