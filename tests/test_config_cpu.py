"""CPU: the reference's configuration file / command line without hydra (oryon_b200/config.py) and the configuration-driven mode
of run_test.py (``python run_test.py -cp exp_data/baseline/ dataset.test.name=nocs test.mask=oracle``, reference README), with
the GPU parts replaced by stand-ins."""
import argparse
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import run_test  # noqa: E402
from oryon_b200 import config, synth  # noqa: E402

# a configuration with the keys of the reference's configs/config.yaml that the test path reads (values are this test's)
YAML = """
exp_name : trial
exp_root : {root}/exp_data
exp_tag : Synthetic
use_seed : False
seed : 7
device : cuda
corrs_device: cpu
compute_vsd: False
dataset:
  root : {root}/data
  batch_size : 4
  img_size : [224,224]
  max_corrs: 500
  test:
    name : nocs
    split : cross_scene_test
    obj : all
model:
  use_catseg_ckpt: True
  image_encoder:
    img_size : [192,192]
    vlm: clip
test:
  mask: predicted
  add_description: 'yes'
  src_sampling : 5000
  solver: pointdsc
  n_corrs: ${{dataset.max_corrs}}
  dist_th: 0.25
pretrained:
  pointdsc: pretrained_models/pointdsc
  vocabulary: {vocab}
eval:
  ckpt:
tmp:
  results_out :
"""


def _write_cfg(tmp_path):
    d = tmp_path / "exp_data" / "trial"
    d.mkdir(parents=True)
    (d / "config.yaml").write_text(YAML.format(root=str(tmp_path), vocab=os.path.join(ROOT, "tests", "golden", "bpe_synth_vocab.txt.gz")))
    return str(d)


def test_load_config_interpolation_and_overrides(tmp_path):
    d = _write_cfg(tmp_path)
    cfg = config.load_config(d, overrides=["dataset.test.name=toyl", "test.mask=oracle", "dataset.max_corrs=300", "seed=3", "use_seed=true",
                                           "+extra.flag=null", "model.image_encoder.img_size=[96,96]"])
    assert cfg.dataset.test.name == "toyl" and cfg["test"]["mask"] == "oracle" and cfg.test.add_description == "yes"
    assert cfg.test.n_corrs == 300 and isinstance(cfg.test.n_corrs, int)              # ${dataset.max_corrs} after the override, typed
    assert cfg.use_seed is True and cfg.seed == 3 and cfg.extra.flag is None and cfg.model.image_encoder.img_size == [96, 96]
    assert cfg.eval.ckpt is None and cfg.tmp.results_out is None and cfg.no_such_key is None
    assert config.select(cfg, "dataset.test.split") == "cross_scene_test" and config.select(cfg, "a.b.c", 5) == 5
    assert config.load_config(os.path.join(d, "config.yaml")).dataset.test.name == "nocs"      # a file path works too
    with pytest.raises(ValueError):
        config.apply_overrides(cfg, ["justakey"])
    # the readers and the pipeline helpers take it as `args`
    from oryon_b200 import pipeline
    assert pipeline._get(cfg, "test.n_corrs") == 300
    csv, metrics, copy = pipeline.pred_filenames(cfg, rand_seed=1)
    assert os.path.basename(csv).startswith("toyl_cross_scene_test_all_")


def test_config_driven_run_on_cpu(monkeypatch, tmp_path):
    from oryon_b200 import datasets, net, pipeline
    from oryon_b200.utils.evaluator import format_sym_set
    from test_evaluator_cpu import _OracleBackend
    from test_run_test_cpu import _fake_rows
    d = _write_cfg(tmp_path)
    info = synth.write_nocs_tree(str(tmp_path / "data"), 0)
    made = {}

    class FakeOryon:
        def __init__(self, args, device, precision=3, tokenizer=None):
            made["model"] = (args.model.use_catseg_ckpt, device, precision)
            self.tokenizer = tokenizer

    class FakePipe:
        def __init__(self, args, test_model=False, model=None):
            made["pipe"] = (args.test.mask, args.test.n_corrs, test_model, model is not None)

        def on_test_start(self, pred_path=None, seed=None):
            made["seed"] = seed

        def test_step(self, batch, batch_idx):
            return _fake_rows(list(range(len(batch["instance_id"]))))

        def on_test_end(self):
            pass

    monkeypatch.setattr(net, "Oryon", FakeOryon)
    monkeypatch.setattr(pipeline, "FPM_Pipeline", FakePipe)
    monkeypatch.setattr(datasets.GpuCollate, "__call__", lambda self, data: dict(instance_id=[s[5] for s in data]))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    orig_score = run_test.score_csv

    def score(csv, ds, exp_tag="", compute_vsd=True, failed=None):
        made["score"] = (exp_tag, compute_vsd)
        models, _, symms = ds.get_object_info()
        return orig_score(csv, ds, exp_tag, compute_vsd, pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}))
    monkeypatch.setattr(run_test, "score_csv", score)
    cfg = config.load_config(d, overrides=["test.mask=oracle"])
    opts = argparse.Namespace(precision=3, out=None, workers=2, no_score=False, no_vsd=False)
    r, w = os.pipe()
    run_test.run_config(cfg, opts, 1, 0, 0, torch.device("cpu"), w)
    os.close(w)
    line = json.loads(os.read(r, 1 << 16).decode())
    os.close(r)
    assert made["model"] == (True, "cuda:0", 3) and made["pipe"] == ("oracle", 500, True, True) and made["seed"] == 7      # args.seed whenever it is set (pipeline.py:296-299); use_seed only gates the dataset constructor
    assert made["score"] == ("Synthetic", False) and line["pairs"] == len(info["pairs"]) and line["batch"] == 4
    results = tmp_path / "exp_data" / "trial" / "results"
    names = sorted(os.listdir(results))
    csv = [n for n in names if n.endswith(".csv")][0]
    assert csv.startswith("nocs_cross_scene_test_all_") and csv[:-4] + ".json" in names
    # the copy of the configuration sits where the reference's offline scorer looks for it (compute_metrics.py:56-58)
    assert "config_" + "_".join(csv[:-4].split("_")[-3:]) + ".yaml" in names
    assert line["csv"] == str(results / csv) and len((results / csv).read_text().splitlines()) == len(info["pairs"])


def test_override_scalars_follow_hydra_not_yaml11(tmp_path):
    """``test.add_description=yes`` must stay the STRING the dataset compares against (YAML 1.1 would make it True)."""
    d = _write_cfg(tmp_path)
    for word in ("yes", "no", "on", "off", "wrong", "desconly"):
        cfg = config.load_config(d, overrides=[f"test.add_description={word}"])
        assert cfg.test.add_description == word and isinstance(cfg.test.add_description, str)
    cfg = config.load_config(d, overrides=["dataset.test.split=007", "dataset.test.obj=1:30", "seed=12", "test.dist_th=1e-1", "use_seed=True",
                                           "tmp.results_out=null", "dataset.img_size=[192, 192]", "exp_tag='yes'"])
    assert cfg.dataset.test.split == "007" and cfg.dataset.test.obj == "1:30" and cfg.seed == 12 and cfg.test.dist_th == 0.1
    assert cfg.use_seed is True and cfg.tmp.results_out is None and cfg.dataset.img_size == [192, 192] and cfg.exp_tag == "yes"
    # the prompt really changes with the flag
    from oryon_b200.datasets import _PairSplitDataset
    ds = _PairSplitDataset.__new__(_PairSplitDataset)
    ds.prompt_templates = ["a photo of a {}."]
    item = {"metadata": {"cls_names": ["mug"], "cls_descs": [["red", "blue"]]}}
    ds.add_description = config.load_config(d, overrides=["test.add_description=yes"]).test.add_description
    assert ds.get_item_prompt(item) == ["red mug", "a photo of a red mug."]
    ds.add_description = config.load_config(d, overrides=["test.add_description=no"]).test.add_description
    assert ds.get_item_prompt(item) == ["mug", "a photo of a mug."]


def test_pipeline_seed_and_null_src_sampling():
    """on_test_start seeds with args.seed whenever it is set (pipeline.py:296-299; use_seed does not gate it), and an
    explicit ``test.src_sampling: null`` means NO source subsample (utils/pcd.py:187), not the default 5000."""
    from oryon_b200 import pipeline
    assert pipeline._get({"test": {"src_sampling": None}}, "test.src_sampling", 5000, keep_none=True) is None
    assert pipeline._get({"test": {}}, "test.src_sampling", 5000, keep_none=True) == 5000
    assert pipeline._get({"test": {"src_sampling": None}}, "test.src_sampling", 5000) == 5000
    assert pipeline._get(config.Config({"test": config.Config({"src_sampling": None})}), "test.src_sampling", 5000, keep_none=True) is None
    assert pipeline._get(config.Config({"test": config.Config()}), "test.src_sampling", 5000, keep_none=True) == 5000
    pipe = pipeline.FPM_Pipeline.__new__(pipeline.FPM_Pipeline)
    pipe.pred_file, pipe.args = None, {"seed": 7, "use_seed": False}
    pipe.on_test_start()
    a = torch.rand(3)
    torch.manual_seed(7)
    assert torch.equal(a, torch.rand(3))
    pipe.args = {"seed": None, "use_seed": True}
    pipe.on_test_start()
    a = torch.rand(3)
    torch.manual_seed(1)
    assert torch.equal(a, torch.rand(3))


def test_per_pair_seed_makes_draws_independent_of_sharding():
    from oryon_b200 import pipeline
    pipe = pipeline.FPM_Pipeline.__new__(pipeline.FPM_Pipeline)
    pipe.pred_file, pipe.args = None, {"seed": 3}
    pipe.n_corrs, pipe.src_sampling, pipe.dist_th, pipe.per_pair_seed = 50, 100, 0.25, True
    pipe.on_test_start()
    g = torch.Generator().manual_seed(0)
    dist = torch.rand(6, 400, generator=g) * 0.5
    n_a = [400, 300, 80, 400, 120, 400]
    whole = pipe.draw_rows(dist, n_a, [True] * 6, pair_index=list(range(10, 16)))
    pipe.on_test_start()
    torch.rand(17)                                   # another rank's generator is somewhere else in its stream
    part = pipe.draw_rows(dist[3:], n_a[3:], [True] * 3, pair_index=[13, 14, 15])
    assert torch.equal(whole[3:], part)
    pipe.per_pair_seed = False
    pipe.on_test_start()
    seq = pipe.draw_rows(dist, n_a, [True] * 6, pair_index=list(range(10, 16)))
    pipe.on_test_start()
    assert torch.equal(seq, pipe.draw_rows(dist, n_a, [True] * 6)) and not torch.equal(seq[3:], part)
