"""CPU, gloo: the sharded test loop of run_test.py (the reference's run_test.py:12-42 without Lightning).  The per-batch step
is a stand-in that returns deterministic records, so what is checked is the host logic: every pair is processed exactly once
whatever the world size, short last batches are handled, and rank 0's prediction CSV -- written from the all_gathered rows in
the reference's wire format -- is byte-identical to the single-process one and parses with the evaluator's reader."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import run_test  # noqa: E402


def _fake_rows(idx):
    rows = []
    for i in idx:
        g = torch.Generator().manual_seed(1000 + i)
        pose = torch.eye(4)
        pose[:3, :] = torch.randn(3, 4, generator=g)
        status = ("ok", "ok", "no_corrs", "ok", "invalid_mask")[i % 5]
        if status != "ok":
            pose = torch.eye(4)
        rows.append(dict(status=status, iou_a=float(np.float32(i / 37.0)), iou_q=float(np.float32(0.25 + i / 91.0)), pred_pose_rel=pose))
    return rows


def _worker(rank, world, n_pairs, batch, port, out, seen_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def step(idx):
        assert 1 <= len(idx) <= batch
        seen.extend(idx)
        return _fake_rows(idx)

    res = run_test.run_sharded(n_pairs, batch, step, out_path=out)
    assert res["world"] == world and len(res["records"]) == n_pairs
    with open(os.path.join(seen_dir, f"seen_{rank}.txt"), "w") as fh:
        fh.write(" ".join(map(str, seen)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs,batch", [(23, 4), (8, 32)])
def test_sharded_loop_world2_matches_single_process(tmp_path, n_pairs, batch):
    single = tmp_path / "single.csv"
    res = run_test.run_sharded(n_pairs, batch, _fake_rows, out_path=str(single))
    assert [r["pair_index"] for r in res["records"]] == list(range(n_pairs))
    assert sum(res["status"].values()) == n_pairs and res["status"]["ok"] == sum(i % 5 in (0, 1, 3) for i in range(n_pairs))

    sharded = tmp_path / "sharded.csv"
    port = 29500 + (os.getpid() + 7 * n_pairs) % 2000
    mp.spawn(_worker, args=(2, n_pairs, batch, port, str(sharded), str(tmp_path)), nprocs=2, join=True)
    assert sharded.read_bytes() == single.read_bytes()
    seen = [list(map(int, (tmp_path / f"seen_{r}.txt").read_text().split())) for r in range(2)]
    assert sorted(seen[0] + seen[1]) == list(range(n_pairs)) and abs(len(seen[0]) - len(seen[1])) <= 1


def test_csv_is_the_reference_wire_format(tmp_path):
    from oryon_b200.utils.evaluator import dict_from_preds
    path = tmp_path / "pred.csv"
    run_test.run_sharded(6, 4, _fake_rows, out_path=str(path))
    lines = path.read_text().splitlines()
    assert len(lines) == 6
    for i, line in enumerate(lines):
        id_a, id_q, pose, iou_a, iou_q = line.split(",")
        assert (id_a, id_q) == run_test.pair_ids(i) and len(pose.split(" ")) == 12
        want = _fake_rows([i])[0]
        np.testing.assert_array_equal(np.array(pose.split(" "), dtype=np.float32).reshape(3, 4), want["pred_pose_rel"][:3].numpy())
        assert np.float32(iou_a) == np.float32(want["iou_a"]) and np.float32(iou_q) == np.float32(want["iou_q"])
    preds, ious_a, ious_q, iou_present = dict_from_preds(str(path))      # the offline scorer's reader (compute_metrics.py:14-49)
    assert iou_present and len(preds) == len(ious_a) == len(ious_q) == 6
    np.testing.assert_allclose(preds["0001_000003_0002_000003_1"], _fake_rows([3])[0]["pred_pose_rel"][:3].numpy(), rtol=1e-7)


def test_dataset_mode_ids_and_config_from_a_nocs_tree(tmp_path):
    """--dataset mode: the reader is configured from the command-line options and the CSV ids are the frames' own ids."""
    import argparse
    from oryon_b200 import synth
    from oryon_b200.datasets import NOCSDataset
    info = synth.write_nocs_tree(str(tmp_path), 0)
    opts = argparse.Namespace(root=str(tmp_path), dataset=info["name"], split=info["split"], obj="all", mask="predicted", add_description="yes")
    ds = NOCSDataset(run_test.dataset_args(opts, "cuda:0"), eval=True)
    assert len(ds) == len(info["pairs"])
    for i, (sa, ia, sq, iq, obj) in enumerate(info["pairs"]):
        id_a, id_q = run_test.dataset_pair_ids(ds, i)
        item_a, item_q, *_ = ds[i]
        assert (id_a, id_q) == (item_a["instance_id"], item_q["instance_id"]) == (f"{sa} {ia} {obj}", f"{sq} {iq} {obj}")
    path = tmp_path / "pred.csv"
    run_test.run_sharded(len(ds), 4, _fake_rows, out_path=str(path), id_fn=lambda i: run_test.dataset_pair_ids(ds, i))
    from oryon_b200.utils.evaluator import dict_from_preds
    preds, *_ = dict_from_preds(str(path))
    assert "1_0_2_0_mug_synth_a" in preds and len(preds) == len(ds)
    # --score: rank 0 scores the gathered CSV over the same dataset (the reference's on_test_end outputs); oracle error backend on the CPU
    from oryon_b200.utils.evaluator import format_sym_set
    from test_evaluator_cpu import _OracleBackend
    models, _, symms = ds.get_object_info()
    ev = run_test.score_csv(str(path), ds, "loop", compute_vsd=False, pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}))
    assert len(ev.metrics["instance_id"]) == len(ds) and (tmp_path / "pred.json").exists()
    assert sum(ev.counts["Missing segm"]) == 1                      # the tree's pair without correspondences


def test_pipeline_test_side_surface(tmp_path):
    """get_dataset / get_pred_filename / get_test_dataloader (pipeline.py:84-98, :474-488, :533-549) as module-level helpers (the
    class itself needs a GPU): dataset selection by name, the file naming the offline scorer relies on, consecutive batches."""
    import datetime
    from oryon_b200 import pipeline, synth
    from oryon_b200.datasets import NOCSDataset, TOYLDataset
    nocs, toyl = synth.write_nocs_tree(str(tmp_path), 0), synth.write_toyl_tree(str(tmp_path), 0)
    args = dict(device="cuda:0", tmp=dict(results_out=str(tmp_path)),
                dataset=dict(root=str(tmp_path), max_corrs=500, img_size=[224, 224], batch_size=4, test=dict(name="nocs", split=nocs["split"], obj="all")),
                test=dict(mask="oracle", add_description="yes"))
    ds = pipeline.get_dataset(args, eval=True)
    assert isinstance(ds, NOCSDataset) and len(ds) == len(nocs["pairs"])
    args["dataset"]["test"]["name"] = "toyl"
    assert isinstance(pipeline.get_dataset(args, eval=True), TOYLDataset)
    args["dataset"]["test"]["name"] = "linemod"
    with pytest.raises(RuntimeError, match="Dataset linemod not supported"):
        pipeline.get_dataset(args, eval=True)
    args["dataset"]["test"]["name"] = "nocs"
    csv, metrics, cfg = pipeline.pred_filenames(args, now=datetime.datetime(2024, 3, 9, 7, 5), rand_seed=42)
    assert os.path.basename(csv) == "nocs_cross_scene_test_all_09032024_0705_42.csv" and metrics == csv[:-4] + ".json"
    assert os.path.basename(cfg) == "config_09032024_0705_42.yaml"
    # the reference scorer rebuilds the configuration name from the last three '_' fields of the CSV name (compute_metrics.py:56-57)
    assert "config_" + "_".join(os.path.splitext(os.path.basename(csv))[0].split("_")[-3:]) + ".yaml" == os.path.basename(cfg)
    loader = pipeline.TestLoader(ds, 4, collate=lambda samples: [s[5] for s in samples])        # stand-in collate: the pair ids
    batches = list(loader)
    assert len(loader) == len(batches) == 2 and [len(b) for b in batches] == [4, 2]
    assert [i for b in batches for i in b] == [ds[i][5] for i in range(len(ds))]
    assert [len(b) for b in pipeline.TestLoader(ds, 4, indices=[3, 4, 5], collate=lambda s: s)] == [3]
    # decoding threads: same batches, same order, reader errors surface at the batch they belong to
    want = list(pipeline.TestLoader(ds, 4, collate=lambda samples: [s[5] for s in samples], workers=0))
    assert list(pipeline.TestLoader(ds, 4, collate=lambda samples: [s[5] for s in samples], workers=3, prefetch=1)) == want == batches

    class Broken:
        def __len__(self):
            return 5

        def __getitem__(self, i):
            if i == 3:
                raise OSError("unreadable frame")
            return i
    got = []
    with pytest.raises(OSError, match="unreadable frame"):
        for b in pipeline.TestLoader(Broken(), 2, collate=lambda s: s, workers=2):
            got.append(b)
    assert got == [[0, 1]]


def test_dataset_mode_glue_on_cpu(monkeypatch, tmp_path):
    """``run_test.run_dataset`` end to end with the GPU parts replaced by stand-ins (pipeline, collate, synchronize; oracle error
    backend for ``--score``): reader -> threaded loader -> step per batch -> gathered CSV with the frames' ids -> scorer JSON ->
    the one result line.  What is NOT covered here is what the stand-ins replace; tools/gpu_dataset_mode.py runs the real thing."""
    import argparse
    import json
    from oryon_b200 import datasets, synth
    from oryon_b200.utils.evaluator import format_sym_set
    from test_evaluator_cpu import _OracleBackend
    info = synth.write_nocs_tree(str(tmp_path), 0)
    seen = []

    class FakePipe:
        def on_test_start(self, pred_path=None, seed=None):
            self.seed = seed

        def test_step(self, batch, batch_idx):
            seen.append(list(batch["instance_id"]))
            return _fake_rows(list(range(len(batch["instance_id"]))))

        def on_test_end(self):
            pass

    monkeypatch.setattr(run_test, "build_pipeline", lambda *a, **k: (FakePipe(), argparse.Namespace(tokenizer=object())))
    monkeypatch.setattr(datasets.GpuCollate, "__call__", lambda self, data: dict(instance_id=[s[5] for s in data]))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    orig_score = run_test.score_csv

    def score(csv, ds, exp_tag="", compute_vsd=True, failed=None):
        models, _, symms = ds.get_object_info()
        scored["failed"] = failed
        return orig_score(csv, ds, exp_tag, compute_vsd, pose_errors=_OracleBackend(models, {k: format_sym_set(s) for k, s in symms.items()}),
                          failed=failed)
    scored = {}
    monkeypatch.setattr(run_test, "score_csv", score)
    out = tmp_path / "pred.csv"
    args = argparse.Namespace(dataset=info["name"], dataset_type=None, root=str(tmp_path), split=info["split"], obj="all", mask="oracle",
                              add_description="yes", batch=4, precision=3, out=str(out), seed=1, workers=2, score=True, no_vsd=True)
    r, w = os.pipe()
    run_test.run_dataset(args, 1, 0, 0, torch.device("cpu"), w)
    os.close(w)
    line = json.loads(os.read(r, 1 << 16).decode())
    os.close(r)
    ids = ["_".join(str(e) for e in (sa, ia, sq, iq, obj)) for sa, ia, sq, iq, obj in info["pairs"]]
    assert seen == [ids[:4], ids[4:]] and line["pairs"] == len(ids) and line["status"]["ok"] + line["status"]["no_corrs"] + line["status"]["invalid_mask"] == len(ids)
    csv_lines = out.read_text().splitlines()
    assert [l.split(",")[0] for l in csv_lines] == [f"{sa} {ia} {obj}" for sa, ia, _, _, obj in info["pairs"]]
    metrics = json.load(open(tmp_path / "pred.json"))
    # in-loop failures (status != ok) are registered as failures, like the reference's on_test_end state (pipeline.py:335-342),
    # on top of the pair the dataset itself marks invalid
    n_failed_in_loop = line["status"]["no_corrs"] + line["status"]["invalid_mask"]
    assert len(scored["failed"]) == n_failed_in_loop and scored["failed"] <= set(ids)
    invalid_by_dataset = 1
    assert metrics["instance_id"] == ids
    assert invalid_by_dataset <= sum(metrics["Missing segm"]) <= invalid_by_dataset + n_failed_in_loop
    missing = {i for i, m in zip(ids, metrics["Missing segm"]) if m}
    assert scored["failed"] <= missing


@pytest.mark.parametrize("n_pairs,batch", [(23, 4), (8, 32), (4, 4)])
def test_sharded_loop_with_a_pipelined_step(tmp_path, n_pairs, batch):
    """A pipelined step (``test.pipelined``) hands back the records of the batch BEFORE the one it was given and ``flush`` the last
    ones: the loop must still attribute every record to its own pair and write the same CSV as the direct loop; without a flush
    function the loop refuses to drop a batch."""
    direct = tmp_path / "direct.csv"
    run_test.run_sharded(n_pairs, batch, _fake_rows, out_path=str(direct))
    held = []

    def step(idx):
        held.append(_fake_rows(idx))
        return held.pop(0) if len(held) > 1 else []

    def flush():
        return held.pop(0) if held else []

    piped = tmp_path / "piped.csv"
    res = run_test.run_sharded(n_pairs, batch, step, out_path=str(piped), flush_fn=flush)
    assert piped.read_bytes() == direct.read_bytes() and [r["pair_index"] for r in res["records"]] == list(range(n_pairs))
    held.clear()
    with pytest.raises(RuntimeError, match="never returned"):
        run_test.run_sharded(n_pairs, batch, step, out_path=str(piped))
