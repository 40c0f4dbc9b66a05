#!/usr/bin/env python
"""Test-loop entry point: the B200 counterpart of the reference's ``run_test.py`` (:12-42) without hydra / Lightning.

    python run_test.py --pairs 2000 --batch 32 --out predictions.csv
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 run_test.py --pairs 2000

What the reference does in one process -- ``Trainer.test(system, test_data)``: ``on_test_start``, ``test_step`` per batch,
``on_test_end`` (pipeline.py:296-370) -- is done here on every rank for its own contiguous share of the pair list
(``sharding.shard_pairs``: no padding, no duplicates), with no collective on the data path; at the end ONE ``all_gather`` of
16-float result rows (``sharding.gather_rows``; NCCL over NVLink on GPUs) brings every pair's pose / IoUs / status to all
ranks and rank 0 writes the single prediction CSV in pair order, in the reference's wire format
(``pipeline.format_pred_line``; scripts/evaluation/compute_metrics.py:14-49 reads it back).  The reference itself has no
such gather: under several GPUs it would leave one partial CSV per process.

No dataset, checkpoint or BPE vocabulary exists offline, so the pairs are synthetic (``synth.synthetic_batch``: 224x224 RGB,
480x640 depth related by a planted rigid motion, NOCS intrinsics) and the weights seeded random ones of the reference's
architecture; with real data the same loop runs over ``GpuCollate`` batches.  Prints one JSON line
(rank 0): pairs, pairs/s (barrier + synchronize on both sides, max over ranks), status counts.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from typing import Callable, Dict, List, Optional, Sequence

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oryon_b200 import sharding  # noqa: E402


def pair_ids(pair_index: int) -> tuple:
    """Instance ids ``'scene image object'`` of a synthetic pair, in the form the offline scorer splits
    (scripts/evaluation/compute_metrics.py:30-33; the reference's ids come from instance_list.txt, datasets.py:421-442)."""
    return f"0001 {pair_index:06d} 1", f"0002 {pair_index:06d} 1"


def run_sharded(n_pairs: int, batch_size: int, step_fn: Callable[[Sequence[int]], List[dict]], *, out_path: Optional[str] = None,
                device: Optional[torch.device] = None, sync: Optional[Callable[[], None]] = None,
                id_fn: Callable[[int], tuple] = None, flush_fn: Optional[Callable[[], List[dict]]] = None) -> Dict:
    """The sharded test loop.  ``step_fn(pair_indices)`` returns one record per pair with the keys ``status``, ``iou_a``,
    ``iou_q``, ``pred_pose_rel`` (``FPM_Pipeline.test_step``'s records).  Returns, on every rank, the gathered table in pair
    order, the status counts and the loop time (max over ranks); rank 0 also writes ``out_path`` (``id_fn(pair_index)`` gives
    the two ``'scene image object'`` ids of a CSV line; default: synthetic ids).  A pipelined step (``test.pipelined``) returns
    the records of an EARLIER batch (or none); ``flush_fn`` (``FPM_Pipeline.flush``) then delivers the last ones: records are
    matched to pair indices in submission order."""
    from oryon_b200.pipeline import format_pred_line
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = sharding.shard_pairs(n_pairs, rank, world)

    def barrier():
        if sync is not None:
            sync()
        if world > 1:
            dist.barrier()

    local, submitted = [], []

    def take(rows):
        if not rows:
            return
        if not submitted or len(rows) != len(submitted[0]):
            raise RuntimeError(f"step_fn returned {len(rows)} records for {len(submitted[0]) if submitted else 0} pairs")
        local.append(sharding.encode_rows(submitted.pop(0), rows))

    barrier()
    t0 = time.perf_counter()
    for s in range(mine.start, mine.stop, batch_size):
        idx = list(range(s, min(s + batch_size, mine.stop)))
        submitted.append(idx)
        take(step_fn(idx))
    if flush_fn is not None:
        take(flush_fn())
    if submitted:
        raise RuntimeError(f"{len(submitted)} batches were submitted and never returned (a pipelined step needs flush_fn)")
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
    local_t = torch.cat(local) if local else torch.zeros(0, sharding.ROW_FLOATS, dtype=torch.float64)
    if device is not None:
        local_t, dt = local_t.to(device), dt.to(device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    table = sharding.gather_rows(local_t, n_pairs).cpu()       # the only collective that carries results
    records = sharding.decode_rows(table)
    if rank == 0 and out_path is not None:
        with open(out_path, "w") as fh:
            for r in records:
                id_a, id_q = (id_fn or pair_ids)(r["pair_index"])
                fh.write(format_pred_line(id_a, id_q, np.float32(r["iou_a"]), np.float32(r["iou_q"]), r["pred_pose_rel"].numpy()))
    counts = {k: sum(r["status"] == k for r in records) for k in sharding.STATUS}
    return dict(table=table, records=records, status=counts, seconds=float(dt.item()), world=world, rank=rank)


def build_pipeline(local_rank: int, precision: int, opts=None, evaluator=None):
    """``FPM_Pipeline(args, test_model=True)`` (run_test.py:15).  Weights: the files named on the command line (the reference's
    ``pretrained_models/`` set and its Lightning checkpoint, run_test.py:42, assembled in the reference's load order by
    ``oryon_b200.checkpoint``), otherwise seeded random weights of the reference's architecture."""
    from oryon_b200 import checkpoint as ck, synth, synth_backbone as sb
    from oryon_b200.net import Oryon
    from oryon_b200.pipeline import FPM_Pipeline
    from oryon_b200.utils.pointdsc.init import PointDSCSolver, get_pointdsc_solver
    cfg = synth.POINTDSC_DEFAULT_CFG
    dev = f"cuda:{local_rank}"
    tokenizer = None
    if opts is not None and opts.bpe:
        from oryon_b200.models.tokenizer import SimpleTokenizer
        tokenizer = SimpleTokenizer(opts.bpe)
    if opts is not None and (opts.clip or opts.ckpt):
        sd = ck.assemble_state_dict(ck.clip_state_dict(opts.clip) if opts.clip else None, ck.swin_state_dict(opts.swin) if opts.swin else None,
                                    torch.load(opts.catseg, map_location="cpu")["model"] if opts.catseg else None,
                                    ck.lightning_model_state_dict(opts.ckpt) if opts.ckpt else None)
    else:
        sd = sb.oryon_state_dict(11)
    model = Oryon(None, dev, state_dict=sd, precision=precision, tokenizer=tokenizer)
    if opts is not None and opts.pointdsc:
        solver = get_pointdsc_solver(opts.pointdsc, dev)
    else:
        solver = PointDSCSolver(synth.pointdsc_state_dict(300), in_dim=cfg["in_dim"], num_layers=cfg["num_layers"],
                                num_channels=cfg["num_channels"], num_iterations=cfg["num_iterations"], ratio=cfg["ratio"],
                                sigma_d=cfg["sigma_d"], k=cfg["k"], nms_radius=cfg["inlier_threshold"], device=dev)
    mask = opts.mask if opts is not None else "oracle"
    args = dict(device=dev, corrs_device="cpu", dataset=dict(img_size=[224, 224], max_corrs=500),
                model=dict(image_encoder=dict(img_size=[192, 192])),
                test=dict(mask=mask, src_sampling=5000, solver="pointdsc", n_corrs=500, dist_th=0.25, mask_threshold=0.5,
                          per_pair_seed=bool(opts is not None and getattr(opts, "per_pair_seed", False)),
                          pipelined=not bool(opts is not None and getattr(opts, "no_pipeline", False))))
    return FPM_Pipeline(args, test_model=True, model=model, pointdsc_solver=solver, evaluator=evaluator), model


def dataset_args(opts, device: str) -> dict:
    """The keys of configs/config.yaml the test-time reader consumes (datasets.py:371-392)."""
    return dict(device=device, dataset=dict(root=opts.root, max_corrs=500, img_size=[224, 224],
                                            test=dict(name=opts.dataset, split=opts.split, obj=opts.obj)),
                test=dict(mask=opts.mask, add_description=opts.add_description))


def dataset_pair_ids(ds, pair_index: int) -> tuple:
    """``'scene image object'`` ids of both frames of a pair, as ``get_item_data`` names them (utils/data/nocs.py:265,
    utils/data/toyl.py:202)."""
    return ds.frame_ids(pair_index)


def pair_instance_id(ds, pair_index: int) -> str:
    """The pair id the dataset gives pair ``pair_index`` (datasets.py:448: ``<scene_a>_<img_a>_<scene_q>_<img_q>_<obj>``)."""
    inst = ds.instances[pair_index]
    return f"{inst[1]}_{inst[2]}_{inst[3]}_{inst[4]}_{inst[-1]}"


def score_csv(csv_path: str, ds, exp_tag: str = "", compute_vsd: bool = True, pose_errors=None, failed=None):
    """Summary, metrics JSON and LaTeX row (pipeline.py:357-370) computed on rank 0 from the gathered prediction CSV by the
    offline scorer (scripts/evaluation/compute_metrics.py) -- every pair exactly once, whatever the number of ranks.  With
    ``failed`` = the ids of the pairs whose in-loop status was not 'ok', those are registered through ``register_test_failure``
    as the reference's test loop does (pipeline.py:335-342: every metric of the pair zero, 'Missing segm' 1), so the outputs are
    what the reference's ``on_test_end`` leaves behind; with ``failed=None`` they follow the OFFLINE scorer's semantics (validity
    from the dataset's oracle mask only, identity pose scored for in-loop failures), which differ whenever a pair fails with
    predicted masks.  The JSON goes next to the CSV, summary and LaTeX row to stdout (= stderr while the loop owns the result
    line)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("oryon_compute_metrics", os.path.join(ROOT, "scripts", "evaluation", "compute_metrics.py"))
    scorer = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scorer)
    return scorer.compute_metrics(csv_path, ds, exp_tag, compute_vsd, True, None, pose_errors=pose_errors, failed=failed)


def run_dataset(args, world: int, rank: int, local: int, dev: torch.device, real_stdout: int) -> None:
    """The test loop over a mounted dataset in the reference's NOCS or TOYL layout: ``NOCSDataset`` / ``TOYLDataset`` samples -> ``GpuCollate`` (decode on
    the host, resize / normalise on the GPU) -> ``test_step``.  Rank 0 writes the prediction CSV; scoring it is the offline
    scorer's job (``oryon_b200.utils.evaluator`` / the reference's scripts/evaluation/compute_metrics.py read it back)."""
    from oryon_b200.datasets import NOCSDataset, TOYLDataset
    cls = {"nocs": NOCSDataset, "toyl": TOYLDataset}[args.dataset_type or ("toyl" if args.dataset.lower().startswith("toyl") else "nocs")]
    ds = cls(dataset_args(args, f"cuda:{local}"), eval=True)
    if len(ds) == 0:
        raise SystemExit(f"run_test.py: no pair of split {args.split!r} matches object split {args.obj!r}")
    pipe, model = build_pipeline(local, args.precision, args)
    if model.tokenizer is None:
        raise SystemExit("run_test.py: text prompts need the CLIP BPE vocabulary (--bpe)")

    dataset_loop(ds, pipe, batch=args.batch, seed=args.seed, workers=args.workers, out=args.out, score=args.score, compute_vsd=not args.no_vsd,
                 label=f"{args.dataset}/{args.split}/{args.obj}", exp_tag=f"{args.dataset} {args.split} {args.obj} ({args.mask})",
                 precision=args.precision, world=world, rank=rank, dev=dev, real_stdout=real_stdout)


def dataset_loop(ds, pipe, *, batch: int, seed: int, workers: int, out: Optional[str], score: bool, compute_vsd: bool, label: str, exp_tag: str,
                 precision: int, world: int, rank: int, dev: torch.device, real_stdout: int) -> Dict:
    """This rank's share of ``ds`` through ``pipe.test_step``, batches decoded on worker threads ahead of the GPU (the reference's
    DataLoader has 8 worker processes), the gathered CSV written by rank 0, optionally scored there; one JSON line on rank 0."""
    from oryon_b200.pipeline import TestLoader
    batches = iter(TestLoader(ds, batch, list(sharding.shard_pairs(len(ds), rank, world)), workers=workers))

    def step(idx: Sequence[int]) -> List[dict]:
        b = next(batches)
        if len(b["instance_id"]) != len(idx):
            raise RuntimeError("run_test.py: loader and loop disagree on the batch boundaries")
        b["pair_index"] = list(idx)
        return pipe.test_step(b, idx[0] // batch)

    pipe.on_test_start(seed=seed)
    res = run_sharded(len(ds), batch, step, out_path=out, device=dev, sync=torch.cuda.synchronize, id_fn=lambda i: dataset_pair_ids(ds, i),
                      flush_fn=getattr(pipe, "flush", None))
    pipe.on_test_end()
    if rank == 0 and score and out is not None:
        failed = {pair_instance_id(ds, r["pair_index"]) for r in res["records"] if r["status"] != "ok"}
        score_csv(out, ds, exp_tag=exp_tag, compute_vsd=compute_vsd, failed=failed)
    if rank == 0:
        line = json.dumps({"metric": "image-pairs/sec (whole test loop, decode included)", "value": len(ds) / res["seconds"], "unit": "pairs/s",
                           "n_gpus": world, "pairs": len(ds), "batch": batch, "seconds": res["seconds"], "status": res["status"],
                           "gemm_precision": precision, "data": label, "csv": out})
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return res


def config_run_files(cfg, rank: int = 0, now=None, rand_seed: Optional[int] = None) -> tuple:
    """Result files of a configuration-driven run, named as ``FPM_Pipeline.get_pred_filename`` names them (pipeline.py:474-488)
    under ``tmp.results_out`` (default ``<exp_root>/<exp_name>/results``, utils/misc.py:385), plus the copy of the configuration
    the reference's offline scorer looks for next to the CSV (compute_metrics.py:56-58).  Written by rank 0."""
    import yaml
    from oryon_b200.config import select
    from oryon_b200.pipeline import pred_filenames
    if select(cfg, "tmp.results_out") is None:
        cfg.setdefault("tmp", type(cfg)())["results_out"] = os.path.join(str(select(cfg, "exp_root", "exp_data")), str(select(cfg, "exp_name", "baseline")), "results")
    csv, metrics, cfg_copy = pred_filenames(cfg, now=now, rand_seed=rand_seed)
    if rank == 0:
        os.makedirs(os.path.dirname(csv), exist_ok=True)
        with open(cfg_copy, "w") as f:
            yaml.safe_dump(json.loads(json.dumps(cfg)), f)
    return csv, metrics, cfg_copy


def run_config(cfg, opts, world: int, rank: int, local: int, dev: torch.device, real_stdout: int) -> None:
    """``python run_test.py -cp exp_data/baseline/ dataset.test.name=nocs test.mask=oracle`` (the reference's command line): the
    whole run is described by the configuration file -- dataset, masks, solver, seeds, ``pretrained.*``, ``eval.ckpt`` -- and the
    weights are read where a reference installation keeps them (``Oryon(args, device)``, oryon_b200/checkpoint.py)."""
    from oryon_b200 import pipeline
    from oryon_b200.config import select
    from oryon_b200.net import Oryon
    cfg["device"] = f"cuda:{local}"
    ds = pipeline.get_dataset(cfg, eval=True)
    if len(ds) == 0:
        raise SystemExit("run_test.py: the pair split is empty for this object split")
    tokenizer = None
    vocab = select(cfg, "pretrained.vocabulary")
    if vocab and os.path.exists(vocab):
        from oryon_b200.models.tokenizer import SimpleTokenizer
        tokenizer = SimpleTokenizer(vocab)
    model = Oryon(cfg, cfg["device"], precision=opts.precision, tokenizer=tokenizer)
    if model.tokenizer is None:
        raise SystemExit(f"run_test.py: CLIP BPE vocabulary not found (pretrained.vocabulary = {vocab})")
    if select(cfg, "test.pipelined") is None and isinstance(cfg.get("test"), dict):
        cfg["test"]["pipelined"] = not getattr(opts, "no_pipeline", False)      # tails run under the next batch's network pass
    pipe = pipeline.FPM_Pipeline(cfg, test_model=True, model=model)
    out = opts.out
    if out is None:
        # every rank derives the same names: the stamp comes from rank 0's clock only through the CSV it alone writes
        out = config_run_files(cfg, rank)[0]
    seed = int(select(cfg, "seed")) if select(cfg, "seed") is not None else 1      # pipeline.py:296-299: args.seed whenever it is set
    name, split, obj = select(cfg, "dataset.test.name"), select(cfg, "dataset.test.split"), select(cfg, "dataset.test.obj")
    dataset_loop(ds, pipe, batch=int(select(cfg, "dataset.batch_size", 32)), seed=seed, workers=opts.workers, out=out, score=not opts.no_score,
                 compute_vsd=bool(select(cfg, "compute_vsd", True)) and not opts.no_vsd, label=f"{name}/{split}/{obj}",
                 exp_tag=str(select(cfg, "exp_tag", "")), precision=opts.precision, world=world, rank=rank, dev=dev, real_stdout=real_stdout)


def main(argv=None):
    # one JSON line on stdout: library banners (NCCL prints its version from C code) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--pairs", type=int, default=2000, help="size of the test split (NOCS / TOYL: 2000 pairs)")
    ap.add_argument("--batch", type=int, default=32, help="configs/config.yaml:17")
    ap.add_argument("--out", default=None, help="prediction CSV written by rank 0")
    ap.add_argument("--precision", type=int, default=2, choices=[1, 2, 3],
                    help="network GEMM precision: 3 = three fp16 products everywhere, 2 = fp8 cross terms in the CLIP vision linear layers (both "
                         "hold the 1e-3 gate), 1 = one product")
    ap.add_argument("--distinct-batches", type=int, default=4, help="synthetic batches generated up front and cycled")
    ap.add_argument("--seed", type=int, default=1, help="on_test_start seed (utils/misc.py:186-196)")
    ap.add_argument("--dataset", default=None, help="name of a mounted dataset in the reference's NOCS layout (args.dataset.test.name), e.g. nocs; "
                                                    "default: synthetic pairs")
    ap.add_argument("--dataset-type", default=None, choices=["nocs", "toyl"], help="on-disk layout (default: from the dataset name)")
    ap.add_argument("--root", default="data", help="args.dataset.root")
    ap.add_argument("--split", default="cross_scene_test", help="args.dataset.test.split (a directory under fixed_split/)")
    ap.add_argument("--obj", default="all", help="args.dataset.test.obj (a key of object_splits.json)")
    ap.add_argument("--mask", default="oracle", choices=["oracle", "predicted", "ovseg", "san", "oryon"], help="args.test.mask")
    ap.add_argument("--add-description", default="yes", choices=["yes", "no", "wrong", "desconly"], help="args.test.add_description")
    ap.add_argument("--bpe", default=None, help="CLIP BPE vocabulary (bpe_simple_vocab_16e6.txt.gz); needed for text prompts")
    ap.add_argument("--clip", default=None, help="OpenAI CLIP ViT-L/14@336 TorchScript archive")
    ap.add_argument("--swin", default=None, help="torchvision swin_b weights")
    ap.add_argument("--catseg", default=None, help="CATSeg checkpoint (pretrained_models/catseg.pth)")
    ap.add_argument("--ckpt", default=None, help="the reference's Lightning checkpoint (args.eval.ckpt)")
    ap.add_argument("-cp", "--config-path", default=None, help="the reference's way: folder (or file) of the hydra configuration, e.g. exp_data/baseline/")
    ap.add_argument("-cn", "--config-name", default="config")
    ap.add_argument("overrides", nargs="*", help="with -cp: dotted overrides, e.g. dataset.test.name=nocs test.mask=oracle")
    ap.add_argument("--no-score", action="store_true", help="with -cp: do not score the CSV at the end")
    ap.add_argument("--torch-threads", type=int, default=2, help="intra-op threads of the host-side torch ops (the per-pair multinomial draws on <= 36 864 "
                                                                  "weights): with the decoding threads busy, a full OpenMP team per tiny op costs milliseconds; "
                                                                  "ignored when OMP_NUM_THREADS is set (torchrun sets it to 1)")
    ap.add_argument("--workers", type=int, default=8, help="dataset mode: decoding threads of the loader (pipeline.py:545 num_workers=8); 0 = in line")
    ap.add_argument("--score", action="store_true", help="dataset mode: rank 0 scores the gathered CSV at the end (metrics JSON next to --out, LaTeX row)")
    ap.add_argument("--no-vsd", action="store_true", help="with --score: skip VSD / AR")
    ap.add_argument("--per-pair-seed", action="store_true", help="re-seed the draw generators from (seed, global pair index) before every pair: "
                                                                 "the prediction rows then do not depend on the number of ranks (SURVEY.md 8e); "
                                                                 "default: the reference's single sequential draw stream per rank")
    ap.add_argument("--no-pipeline", action="store_true", help="run every batch's post-network tail before the next network pass instead of under it")
    ap.add_argument("--pointdsc", default=None, help="PointDSC snapshot directory (args.pretrained.pointdsc)")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("run_test.py: no CUDA device; the product path has no CPU fallback")
    if "OMP_NUM_THREADS" not in os.environ:
        torch.set_num_threads(max(1, args.torch_threads))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from oryon_b200 import synth
    if args.config_path is not None:
        from oryon_b200.config import load_config
        return run_config(load_config(args.config_path, args.config_name, args.overrides), args, world, rank, local, dev, real_stdout)
    if args.overrides:
        raise SystemExit(f"run_test.py: overrides {args.overrides} need a configuration (-cp)")
    if args.dataset is not None:
        return run_dataset(args, world, rank, local, dev, real_stdout)
    pipe, model = build_pipeline(local, args.precision, args)
    batches = []
    for k in range(args.distinct_batches):
        b = synth.synthetic_batch(100 + k, args.batch)
        emb = model.encode_tokens(b.pop("prompt_tokens")[0].cuda())[None]      # one prompt set per batch: cached as the benchmark allows
        for key in ("anchor", "query"):
            b[key]["rgb"] = b[key]["rgb"].pin_memory()
            b[key]["orig_depth"] = torch.stack(b[key]["orig_depth"]).pin_memory()
        b["prompt_emb_one"] = emb
        batches.append(b)

    def take(b: dict, n: int) -> dict:
        """The first n pairs of a generated batch (the last batch of a shard may be short)."""
        out = dict(prompt_emb=b["prompt_emb_one"].expand(n, -1, -1).contiguous(), instance_id=b["instance_id"][:n], cls_id=b["cls_id"][:n])
        for key in ("anchor", "query"):
            out[key] = {k: v[:n] for k, v in b[key].items()}
        return out

    def step(idx: Sequence[int]) -> List[dict]:
        b = take(batches[(idx[0] // args.batch) % len(batches)], len(idx))
        b["pair_index"] = list(idx)
        return pipe.test_step(b, idx[0] // args.batch)

    pipe.on_test_start(seed=args.seed)              # one rank = the reference's draw order; --per-pair-seed: sharding-independent rows
    step(list(range(min(args.batch, args.pairs))))  # warm-up: workspaces, arena, clocks
    pipe.flush()
    pipe.on_test_start(seed=args.seed)
    res = run_sharded(args.pairs, args.batch, step, out_path=args.out, device=dev, sync=torch.cuda.synchronize, flush_fn=pipe.flush)
    pipe.on_test_end()
    if rank == 0:
        line = json.dumps({"metric": "image-pairs/sec (whole test loop)", "value": args.pairs / res["seconds"], "unit": "pairs/s",
                          "n_gpus": world, "pairs": args.pairs, "batch": args.batch, "seconds": res["seconds"], "status": res["status"],
                          "gemm_precision": args.precision, "data": "synthetic", "csv": args.out,
                          "config": {"workload": f"{args.pairs} synthetic pairs: 224x224 RGB -> CLIP ViT-L/14@336 + swin_b + fusion + decoder -> "
                                                 "matching -> lift -> PointDSC; pairs sharded over ranks, one all_gather of result rows"}})
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
