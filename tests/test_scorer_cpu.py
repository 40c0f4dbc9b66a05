"""N3: the offline scorer (scripts/evaluation/compute_metrics.py) against what the reference's own scorer wrote for the same
prediction CSV over the same synthetic TOYL tree at 480 x 640 (``oracle/make_golden_scorer.py`` -> ``tests/golden/scorer_0.json``):
the metrics JSON of ``Evaluator.save`` entry by entry and the LaTeX row, with VSD / AR on.  On the CPU the pose errors come from
the oracles (injected backend); the ``gpu`` test runs the same comparison on the CUDA library, the product path."""
import importlib.util
import json
import os

import numpy as np
import pytest

from oryon_b200 import synth
from oryon_b200.datasets import NOCSDataset, TOYLDataset
from oryon_b200.utils.evaluator import format_sym_set

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "scorer_0.json")))
GOLD["nocs"] = json.load(open(os.path.join(ROOT, "tests", "golden", "scorer_nocs_0.json")))      # NOCS tree, pairs keyed by object name, VSD off

_spec = importlib.util.spec_from_file_location("oryon_compute_metrics", os.path.join(ROOT, "scripts", "evaluation", "compute_metrics.py"))
scorer = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(scorer)


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("toyl_scorer"))
    return d, synth.write_toyl_tree(d, 0, hw=(480, 640))


@pytest.fixture(scope="module")
def nocs_tree(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("nocs_scorer"))
    return d, synth.write_nocs_tree(d, 0)


def _dataset(tree, obj_split):
    d, info = tree
    if info["name"] == "nocs":
        args = dict(device="cuda:0", dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj=obj_split)),
                    test=dict(mask="predicted", add_description="yes"))
        return NOCSDataset(args, eval=True)
    args = dict(device="cuda:0", dataset=dict(root=d, max_corrs=500, img_size=[224, 224], test=dict(name=info["name"], split=info["split"], obj=obj_split)),
                test=dict(mask="predicted", add_description="yes"))
    return TOYLDataset(args, eval=True)


def _score(tree, tag, backend_factory, tmp_path):
    gold = GOLD[tag]
    ds = _dataset(tree, gold["obj"])
    csv = tmp_path / f"toyl_{tag}.csv"
    csv.write_text("".join(gold["csv"]))
    tex = tmp_path / f"{tag}.tex"
    models, _, symms = ds.get_object_info()
    vsd = tag != "nocs"
    ev = scorer.compute_metrics(str(csv), ds, "synthetic nocs" if tag == "nocs" else "synthetic", vsd, False, str(tex),
                                pose_errors=backend_factory(models, symms))
    got = json.load(open(tmp_path / f"toyl_{tag}.json"))
    assert list(got.keys()) == list(gold["metrics"].keys())
    for k, v in gold["metrics"].items():
        if all(isinstance(x, str) for x in v):
            assert got[k] == v, k
        else:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64), np.asarray(v, dtype=np.float64), rtol=1e-9, atol=1e-7, err_msg=k)
    for k in (("VSD", "AR") if vsd else ()) + ("MSSD", "MSPD", "ADD(S)-0.1d"):
        assert [float(x) for x in got[k]] == [float(x) for x in gold["metrics"][k]], k       # thresholded: exact
    assert tex.read_text() == gold["latex"]
    return ev


def _oracle_backend(models, symms):
    from test_evaluator_cpu import _OracleBackend
    return _OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()})


@pytest.mark.parametrize("tag", ["noiou", "iou"])
def test_scorer_matches_reference_scorer(tree, tag, tmp_path):
    ev = _score(tree, tag, _oracle_backend, tmp_path)
    if tag == "noiou":
        assert "1_0_2_2_5" in ev.metrics["instance_id"]              # the pair without correspondences: a failure row


def test_scorer_matches_reference_scorer_nocs(nocs_tree, tmp_path):
    ev = _score(nocs_tree, "nocs", _oracle_backend, tmp_path)
    assert "1_1_2_0_bowl_synth_b" in ev.metrics["instance_id"] and sum(ev.counts["Missing segm"]) == 1


def test_scorer_failure_rows_keep_ious(tree, tmp_path):
    """With IoUs in the CSV and an invalid pair in the split the reference's scorer raises KeyError('iou_a') (its failure
    payload has no IoUs, compute_metrics.py:111-114 / evaluator.py:311); this one registers the failure with the CSV's IoUs."""
    ds = _dataset(tree, "all")
    csv = tmp_path / "toyl_all_iou.csv"
    csv.write_text("".join(GOLD["iou"]["csv"]))
    models, _, symms = ds.get_object_info()
    ev = scorer.compute_metrics(str(csv), ds, "synthetic", False, False, str(tmp_path / "x.tex"), pose_errors=_oracle_backend(models, symms))
    assert len(ev.metrics["instance_id"]) == len(ds) and len(ev.metrics["Anchor IoU"]) == len(ds)


def test_scorer_cli_dataset_detection(tmp_path):
    with pytest.raises(RuntimeError, match="Dataset not supported"):
        scorer.main([str(tmp_path / "preds.csv")])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["noiou", "iou"])
def test_scorer_matches_reference_scorer_cuda(tree, tag, tmp_path):
    from gpu_util import need_gpu
    need_gpu()
    _score(tree, tag, lambda models, symms: None, tmp_path)          # None -> the Evaluator builds CudaPoseErrors
