"""Mirror of the inference part of the reference ``pipeline.py``: ``FPM_Pipeline`` with ``forward`` (:593),
``test_step`` (:306-355), ``is_detection_valid`` (:372-395), ``get_featmap_corrs`` (:397-427), ``get_pose``
(:429-472), ``add_pred_pose`` (:490-497), ``get_dataset`` (:84-98), ``get_pred_filename`` (:474-488) and
``get_test_dataloader`` (:533-549).  Same names, argument meaning, failure rows and CSV wire format.

Every arithmetic step runs in liboryon_b200.so: the network (``oryon_backbone_forward``), mask post-processing
(``oryon_mask_postproc``), matching (``oryon_match_nn``), scaling / lifting (``oryon_corrs_to_pcd``) and
registration (``oryon_pointdsc_pose``).  What stays in torch is what the reference leaves to torch's generator: the
two ``multinomial`` draws per pair, in the reference's order.

``test_step`` is batched the B200 way -- one network pass, one matching call and one registration call for the
whole batch -- yet produces exactly what the reference's per-pair loop produces: the nearest neighbour of an anchor
pixel does not depend on which other anchor pixels were sub-sampled, so all ROI pixels of all pairs are matched in
one launch and the (cheap, ordered) random selections are applied afterwards, pair by pair, in the reference's
draw order.

Lightning / hydra are not dependencies: the class is a plain object with the LightningModule hook names; ``args`` is
any attribute-style mapping with the keys of configs/config.yaml that the inference path reads.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._torch_glue import as_device, ptr, stream_ptr
from .net import Oryon
from .utils import pcd as _pcd
from .utils.pcd import corrs_to_pcd, mask_to_roi, match_nn, nn_correspondences, select_lift_batched, torch_sample_select
from .utils.pointdsc.init import get_pointdsc_pose, get_pointdsc_solver, pointdsc_poses


_ABSENT = object()


def _get(obj, path: str, default=None, *, keep_none: bool = False):
    """``obj.a.b.c`` of a dict / attribute-style configuration.  A missing key gives ``default``; a key that is present with the
    value ``None`` (``src_sampling: null``) gives ``default`` too unless ``keep_none`` -- the reference distinguishes the two where
    ``None`` is a setting of its own (``if subsample_source is not None``, utils/pcd.py:187)."""
    cur = obj
    for key in path.split("."):
        if cur is None or cur is _ABSENT:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key, _ABSENT)
        else:
            cur = getattr(cur, key, _ABSENT)
    if cur is _ABSENT:
        return default
    if cur is None:
        return None if keep_none else default
    return cur


def mask_postproc(logits: Optional[Tensor], gt: Optional[Tensor], size: Tuple[int, int], mask_th: float = 0.5) -> Dict[str, Tensor]:
    """``oryon_mask_postproc``: predicted mask, nearest-resized ground truth, counts and IoU for a batch.
    ``logits [B,1,H,W]`` or ``[B,H,W]`` (or ``None``), ``gt [B,Hg,Wg]`` (or ``None``)."""
    src = logits if logits is not None else gt
    dev = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
    H, W = int(size[0]), int(size[1])
    lg = None
    if logits is not None:
        lg = as_device(logits, dev, torch.float32).reshape(-1, H, W)
    g = None
    if gt is not None:
        g = as_device(gt, dev).to(torch.uint8).contiguous()
    B = lg.shape[0] if lg is not None else g.shape[0]
    out = dict(n_pred=torch.zeros(B, dtype=torch.int32, device=dev), n_gt=torch.zeros(B, dtype=torch.int32, device=dev))
    out["pred"] = torch.empty(B, H, W, dtype=torch.int32, device=dev) if lg is not None else None
    out["gt_resized"] = torch.empty(B, H, W, dtype=torch.int32, device=dev) if g is not None else None
    out["iou"] = torch.empty(B, dtype=torch.float32, device=dev) if (lg is not None and g is not None) else None
    _lib.check(_lib.load().oryon_mask_postproc(
        _lib.handle(dev.index), ptr(lg), B, H, W, float(mask_th), ptr(g), 0 if g is None else g.shape[1], 0 if g is None else g.shape[2],
        ptr(out["pred"]), ptr(out["gt_resized"]), ptr(out["n_pred"]), ptr(out["n_gt"]), ptr(out["iou"]), stream_ptr(dev)))
    return out


def format_pred_line(id_a: str, id_q: str, mask_a_iou, mask_q_iou, pred_pose: np.ndarray) -> str:
    """The prediction-CSV wire format of the reference (pipeline.py:490-497): ``str()`` of the numpy scalars as they are
    (float32 poses and IoUs print their shortest float32 representation)."""
    pose = " ".join([str(n) for n in pred_pose[:3, :].flatten()])
    return ",".join([id_a, id_q, pose, str(mask_a_iou), str(mask_q_iou)]) + "\n"


def get_dataset(args, eval: bool = True):
    """``FPM_Pipeline.get_dataset`` (pipeline.py:84-98): the dataset class named by ``args.dataset.test.name`` (``train.name``
    when not ``eval``).  The test-time readers are here; 'shapenet6d' is the reference's TRAINING set and is not."""
    from .datasets import NOCSDataset, TOYLDataset
    name = _get(args, "dataset.test.name" if eval else "dataset.train.name")
    if name == "nocs":
        return NOCSDataset(args, eval)
    if name == "toyl":
        return TOYLDataset(args, eval)
    if name == "shapenet6d":
        raise NotImplementedError("Dataset shapenet6d is the reference's training set: outside the inference path")
    raise RuntimeError(f"Dataset {name} not supported")


def pred_filenames(args, now=None, rand_seed: Optional[int] = None) -> Tuple[str, str, str]:
    """The three result files of a test run (pipeline.py:474-488): ``<name>_<split>_<obj>_<ddmmYYYY_HHMM>_<rand>.csv`` / ``.json``
    and ``config_<ddmmYYYY_HHMM>_<rand>.yaml`` under ``args.tmp.results_out`` -- the naming the reference's offline scorer
    relies on to find the configuration next to a CSV (scripts/evaluation/compute_metrics.py:56-58).  ``rand`` is drawn from
    numpy's global generator as in the reference."""
    import os
    from datetime import datetime
    stamp = (now or datetime.now()).strftime("%d%m%Y_%H%M")
    rand = int(np.random.randint(0, 1000)) if rand_seed is None else int(rand_seed)
    base = f"{_get(args, 'dataset.test.name')}_{_get(args, 'dataset.test.split')}_{_get(args, 'dataset.test.obj')}_{stamp}_{rand}"
    out = _get(args, "tmp.results_out", ".")
    return os.path.join(out, base + ".csv"), os.path.join(out, base + ".json"), os.path.join(out, f"config_{stamp}_{rand}.yaml")


class TestLoader:
    """What ``get_test_dataloader`` returns in place of ``DataLoader(test_set, batch_size, collate_fn=test_set.collate,
    shuffle=False, num_workers=8)`` (pipeline.py:533-549): batches of consecutive samples in dataset order, the last one short,
    each collated by the dataset's ``GpuCollate`` (resize / normalise on the GPU) in the consumer's thread, on its stream.
    The host part of a sample -- PNG decoding, ~10-20 ms per 480 x 640 frame, what the reference gives its 8 worker processes --
    runs on ``workers`` threads (the decoders release the GIL), ``prefetch`` batches ahead of the consumer, so that the GPU step
    of batch k overlaps the decoding of batches k+1.. .  ``indices`` restricts the loader to a rank's share of the pairs
    (``sharding.shard_pairs``); ``workers=0`` reads in the calling thread.  With the worker threads busy, keep torch's intra-op thread
    count small (``torch.set_num_threads(1..2)``, as ``run_test.py`` does): the loop's own CPU ops are tiny and a full OpenMP team
    per op costs more than the op."""
    __test__ = False            # not a pytest class

    def __init__(self, dataset, batch_size: int, indices: Optional[List[int]] = None, collate=None, workers: int = 8, prefetch: int = 2):
        self.dataset, self.batch_size = dataset, int(batch_size)
        self.indices = list(range(len(dataset))) if indices is None else list(indices)
        self.collate = collate if collate is not None else dataset.collate
        self.workers, self.prefetch = int(workers), max(1, int(prefetch))

    def __len__(self) -> int:
        return (len(self.indices) + self.batch_size - 1) // self.batch_size

    def _chunks(self) -> List[List[int]]:
        return [self.indices[b0:b0 + self.batch_size] for b0 in range(0, len(self.indices), self.batch_size)]

    def __iter__(self):
        chunks = self._chunks()
        if self.workers <= 0:
            for chunk in chunks:
                yield self.collate([self.dataset[i] for i in chunk])
            return
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(self.workers, thread_name_prefix="oryon-decode") as pool:
            pending: deque = deque()
            nxt = 0
            try:
                while nxt < len(chunks) or pending:
                    while nxt < len(chunks) and len(pending) <= self.prefetch:
                        pending.append([pool.submit(self.dataset.__getitem__, i) for i in chunks[nxt]])
                        nxt += 1
                    yield self.collate([f.result() for f in pending.popleft()])      # a reader error surfaces here, in order
            finally:
                for futs in pending:
                    for f in futs:
                        f.cancel()


class FPM_Pipeline:
    def __init__(self, args, test_model: bool = False, *, model: Optional[Oryon] = None, pointdsc_solver=None, evaluator=None):
        self.args = args
        self.test_model = test_model
        self.device = torch.device(_get(args, "device", "cuda"))
        if self.device.type != "cuda":
            raise _lib.OryonError("oryon_b200.pipeline.FPM_Pipeline runs on a CUDA (sm_100) device only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = model if model is not None else Oryon(args, self.device)
        self.corrs_device = _get(args, "corrs_device", "cpu")
        self.solver = _get(args, "test.solver", "pointdsc")
        self.pointdsc_solver = pointdsc_solver
        if self.solver == "pointdsc" and self.pointdsc_solver is None:
            self.pointdsc_solver = get_pointdsc_solver(_get(args, "pretrained.pointdsc"), self.device)
        self.mask_mode = _get(args, "test.mask", "predicted")
        self.mask_th = float(_get(args, "test.mask_threshold", 0.5))
        self.dist_th = float(_get(args, "test.dist_th", 0.25))
        self.n_corrs = int(_get(args, "test.n_corrs", _get(args, "dataset.max_corrs", 500)))
        self.src_sampling = _get(args, "test.src_sampling", 5000, keep_none=True)   # an explicit null = no source subsample
        self.per_pair_seed = bool(_get(args, "test.per_pair_seed", False))          # draws seeded per pair
        self.featmap_size = tuple(_get(args, "model.image_encoder.img_size", (192, 192)))
        self.pred_file = None
        self.rows: List[dict] = []
        self.evaluator = evaluator                                        # oryon_b200.utils.evaluator.Evaluator or None
        self.batched_tail = bool(_get(args, "test.batched_tail", True))   # False: per-pair selection / lifting (any frame sizes)
        self._host_dist: Optional[Tensor] = None
        self._host_rows: Optional[Tensor] = None
        # test.pipelined: the post-network tail of batch k (matching, draws, lifting, registration, rows) runs on a second,
        # high-priority stream UNDER the network pass of batch k+1 (see test_step).  Off by default: test_step then returns the
        # rows of the batch it was given.
        self.pipelined = bool(_get(args, "test.pipelined", False))
        self._pending: Optional[dict] = None
        self._tail_stream: Optional[torch.cuda.Stream] = None

    # ---- LightningModule surface -----------------------------------------------------------------------
    def forward(self, x: dict) -> Dict[str, Tensor]:
        return self.model.forward(x)

    def get_dataset(self, eval: bool = True):
        return get_dataset(self.args, eval)

    def get_pred_filename(self) -> Tuple[str, str]:
        """Paths of the prediction CSV and the metrics JSON (pipeline.py:474-488; the reference returns open file objects and also
        dumps its hydra configuration next to them)."""
        csv, metrics, _ = pred_filenames(self.args)
        return csv, metrics

    def get_test_dataloader(self, indices: Optional[List[int]] = None) -> "TestLoader":
        test_set = self.get_dataset(eval=True)
        print("TESTING on {}, split {}, object split {}. Samples: {}".format(test_set.name, test_set.split, test_set.obj, len(test_set)))
        self.test_dataset = test_set
        return TestLoader(test_set, int(_get(self.args, "dataset.batch_size", 32)), indices)

    def get_valid_dataloader(self, indices: Optional[List[int]] = None) -> "TestLoader":
        """The validation loader reads the same evaluation dataset, in order (pipeline.py:517-531)."""
        valid_set = self.get_dataset(eval=True)
        print("VALIDATING on {}, split {}, object split {}. Samples: {}".format(valid_set.name, valid_set.split, valid_set.obj, len(valid_set)))
        self.valid_dataset = valid_set
        return TestLoader(valid_set, int(_get(self.args, "dataset.batch_size", 32)), indices)

    def on_test_start(self, pred_path: Optional[str] = None, seed: Optional[int] = None):
        """Opens the prediction CSV and seeds numpy / torch as ``set_deterministic_seed`` (utils/misc.py:186-196) with
        ``args.seed`` whenever it is not None, else 1 (pipeline.py:296-299; ``use_seed`` only gates the dataset constructor)."""
        if pred_path is not None:
            self.pred_file = open(pred_path, "w")
        self._pending = None
        if seed is None:
            cfg_seed = _get(self.args, "seed", None)
            seed = int(cfg_seed) if cfg_seed is not None else 1
        self._seed = int(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        self.rows = []

    def flush(self) -> List[dict]:
        """Pipelined mode: finishes the batch whose tail is still outstanding and returns its rows (``[]`` if there is none)."""
        rows: List[dict] = []
        if self._pending is not None:
            pending, self._pending = self._pending, None
            rows = self._back(pending)
        return rows

    def on_test_end(self) -> List[dict]:
        """Closes the prediction CSV (after the last outstanding batch of the pipelined mode, whose rows are returned)."""
        rows = self.flush()
        if self.pred_file is not None:
            self.pred_file.close()
            self.pred_file = None
        return rows

    # ---- a7 -----------------------------------------------------------------------------------------------
    def mask_results(self, batch: dict, outputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """The entries of ``FeatureLoss.forward``'s ``results`` that inference consumes (losses.py:56-60 via :64-141):
        ``mask_a/q`` int ``[B,H,W]``, ``iou_a/q [B]``, ``logits_a/q [B,H,W]``, plus the pixel counts."""
        res = {}
        for v, key in (("a", "anchor"), ("q", "query")):
            gt = batch[key].get("mask") if isinstance(batch.get(key), dict) else None
            mp = mask_postproc(outputs["mask_" + v], gt, self.featmap_size, self.mask_th)
            res["mask_" + v], res["iou_" + v] = mp["pred"], mp["iou"]
            res["logits_" + v] = outputs["mask_" + v].squeeze(1)
            res["n_pred_" + v], res["n_gt_" + v], res["gt_" + v] = mp["n_pred"], mp["n_gt"], mp["gt_resized"]
        return res

    def _masks_and_counts(self, results: dict) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        if self.mask_mode != "predicted":  # external mask, nearest-resized to the feature map (pipeline.py:378-386, :407-412)
            return results["gt_a"], results["gt_q"], results["n_gt_a"], results["n_gt_q"]
        return results["mask_a"], results["mask_q"], results["n_pred_a"], results["n_pred_q"]

    def is_detection_valid(self, results: dict, batch: dict, idx: int) -> bool:
        """True when both views have at least one mask pixel equal to 1 (pipeline.py:372-395)."""
        _, _, na, nq = self._masks_and_counts(results)
        return bool(na[idx].item() > 0) and bool(nq[idx].item() > 0)

    # ---- a8 -----------------------------------------------------------------------------------------------
    def get_featmap_corrs(self, batch: dict, net_output: dict, results: dict, idx: int):
        """Correspondences of pair ``idx`` between the two feature maps (pipeline.py:397-427)."""
        ma, mq, _, _ = self._masks_and_counts(results)
        fa, fq = net_output["featmap_a"][idx], net_output["featmap_q"][idx]
        pred_corrs = nn_correspondences(fa, fq, ma[idx], mq[idx], self.dist_th, self.n_corrs, self.src_sampling, self.corrs_device)
        if pred_corrs is not None:
            ca, cq = pred_corrs[:, :2], pred_corrs[:, 2:]
            pos_a = fa[:, ca[:, 0], ca[:, 1]].transpose(1, 0)
            pos_q = fq[:, cq[:, 0], cq[:, 1]].transpose(1, 0)
        else:
            pos_a, pos_q = None, None
        return pred_corrs, pos_a, pos_q

    # ---- a9-a11 -------------------------------------------------------------------------------------------
    def _lift(self, batch: dict, corrs: Tensor, idx: int) -> Tuple[Tensor, Tensor]:
        depth_a, depth_q = batch["anchor"]["orig_depth"][idx].squeeze(), batch["query"]["orig_depth"][idx].squeeze()
        cam_a, cam_q = batch["anchor"]["camera"][idx].reshape(9), batch["query"]["camera"][idx].reshape(9)
        HA, WA = (int(v) for v in batch["anchor"]["sizes"][idx])
        HQ, WQ = (int(v) for v in batch["query"]["sizes"][idx])
        return corrs_to_pcd(corrs, as_device(depth_a, self.device), as_device(depth_q, self.device), cam_a, cam_q, self.featmap_size,
                            (HA, WA), (HQ, WQ))

    def get_pose(self, batch: dict, corrs: Tensor, idx: int) -> Tensor:
        """3D-3D correspondences of pair ``idx`` and their registration -> ``[4,4]`` float32 CPU (pipeline.py:429-472)."""
        pcd_a, pcd_q = self._lift(batch, corrs, idx)
        if self.solver == "pointdsc":
            pose4 = get_pointdsc_pose(self.pointdsc_solver, pcd_a, pcd_q, self.device)
        else:
            raise RuntimeError(f"Solver {self.solver} not implemented")
        return pose4.to(torch.float32)

    # ---- a12 ----------------------------------------------------------------------------------------------
    def add_pred_pose(self, id_a: str, id_q: str, mask_a_iou, mask_q_iou, pred_pose: np.ndarray):
        """One CSV line ``id_a,id_q,<12 floats>,iou_a,iou_q`` (pipeline.py:490-497; read back by
        scripts/evaluation/compute_metrics.py:14-49)."""
        line = format_pred_line(id_a, id_q, mask_a_iou, mask_q_iou, pred_pose)
        if self.pred_file is not None:
            self.pred_file.write(line)
        return line

    def select_correspondences(self, results: dict, net_output: dict, valid: List[bool],
                               pair_index: Optional[Sequence[int]] = None) -> List[Optional[Tensor]]:
        """Batched ``nn_correspondences`` for every valid pair: one ROI compaction and one matching launch over ALL
        ROI pixels, then per pair, in order, the reference's two draws (utils/pcd.py:187-190, :211) and selections."""
        ma, mq, na, nq = self._masks_and_counts(results)
        fa, fq = net_output["featmap_a"], net_output["featmap_q"]
        B, D, H, W = fa.shape
        roi_a, cnt_a = mask_to_roi(ma)
        roi_q, cnt_q = mask_to_roi(mq)
        n_a, n_q = cnt_a.tolist(), cnt_q.tolist()
        n_a = [n if v else 0 for n, v in zip(n_a, valid)]
        n_q = [n if v else 0 for n, v in zip(n_q, valid)]
        out: List[Optional[Tensor]] = [None] * B
        if not any(valid):
            return out
        idx, dist = match_nn(fa, fq, roi_a, roi_q, n_a, n_q)
        for b in range(B):
            if not valid[b]:
                continue
            self._seed_pair(pair_index, b)
            n1 = n_a[b]
            pix1, pix2 = roi_a[b, :n1], roi_q[b, :n_q[b]]
            ib, db = idx[b, :n1], dist[b, :n1]
            if self.src_sampling is not None and n1 > self.src_sampling:
                sel = torch_sample_select(torch.empty(n1, 0, device=self.corrs_device), int(self.src_sampling)).to(self.device)
                pix1, ib, db = pix1[sel], ib[sel], db[sel]
            ok = torch.nonzero(db < self.dist_th).squeeze(1)
            if ok.shape[0] > 1:
                p1 = pix1[ok].long()
                p2 = pix2[ib[ok].long()].long()
                final = torch.stack((p1 // W, p1 % W, p2 // W, p2 % W), dim=1)
                sel2 = torch_sample_select(torch.empty(final.shape[0], 0, device=self.corrs_device), self.n_corrs).to(self.device)
                out[b] = final[sel2]
        return out

    # ---- batched tail: one host round trip for the draws, one launch for selection + lifting -----------------
    def _stacked_depth(self, view: dict) -> Optional[Tensor]:
        """``orig_depth`` as one ``[B,H,W]`` device tensor, or ``None`` when the frames of the batch differ in size
        (then the per-pair path is used).  A stacked (ideally pinned) tensor from the collate goes up in one copy."""
        d = view["orig_depth"]
        if isinstance(d, Tensor):
            d = d.reshape(d.shape[0], *d.shape[-2:]) if d.dim() == 4 else d
            return as_device(d, self.device)
        frames = [f.squeeze() for f in d]
        if any(f.shape != frames[0].shape or f.dtype != frames[0].dtype for f in frames):
            return None
        out = torch.empty(len(frames), *frames[0].shape, dtype=frames[0].dtype, device=self.device)
        for b, f in enumerate(frames):
            out[b].copy_(f, non_blocking=True)
        return out

    def _uniform_sizes(self, batch: dict, key: str, depth: Optional[Tensor]) -> bool:
        if depth is None:
            return False
        sz = batch[key]["sizes"]
        sz = sz.cpu() if isinstance(sz, Tensor) else torch.as_tensor(sz)
        return bool((sz[:, 0] == depth.shape[1]).all()) and bool((sz[:, 1] == depth.shape[2]).all())

    def _seed_pair(self, pair_index: Optional[Sequence[int]], b: int) -> None:
        """``test.per_pair_seed``: the generators are re-seeded from (run seed, global pair index) before a pair's draws, so a
        pair's correspondences -- hence its CSV line -- do not depend on how the pair list was sharded.  Off by
        default: one rank then reproduces the reference's single sequential draw stream."""
        if self.per_pair_seed and pair_index is not None:
            torch.manual_seed((getattr(self, "_seed", 1) * 1000003 + int(pair_index[b])) & 0x7FFFFFFFFFFFFFFF)

    def draw_rows(self, dist: Tensor, n_a: List[int], valid: List[bool], pair_index: Optional[Sequence[int]] = None) -> Tensor:
        """The two draws of ``nn_correspondences`` (utils/pcd.py:187-190, :211) for every valid pair, in the
        reference's order, on the CPU generator, from the HOST copy of the nearest-neighbour distances.  Returns
        ``int32 [B,n_corrs]`` positions in each pair's anchor ROI list (first entry -1: no correspondences)."""
        B = dist.shape[0]
        rows = torch.full((B, self.n_corrs), -1, dtype=torch.int32)
        for b in range(B):
            if not valid[b]:
                continue
            self._seed_pair(pair_index, b)
            n1 = n_a[b]
            sel = None
            if self.src_sampling is not None and n1 > self.src_sampling:
                sel = torch_sample_select(torch.empty(n1, 0), int(self.src_sampling))
                db = dist[b].index_select(0, sel)
            else:
                db = dist[b, :n1]
            ok = torch.nonzero(db < self.dist_th).squeeze(1)
            if ok.shape[0] > 1:
                r = ok[torch_sample_select(torch.empty(ok.shape[0], 0), self.n_corrs)]
                rows[b] = (sel[r] if sel is not None else r).to(torch.int32)
        return rows

    def _batched_tail(self, batch: dict, outputs: dict, results: dict, valid: List[bool]):
        """Matching -> draws -> selection + lifting -> registration with two host synchronisations per BATCH (the
        distances for the draws, the per-pair point counts) instead of several per pair.  Returns ``(corrs, poses)``
        or ``None`` when the batch does not qualify (frames of different sizes, CUDA-generator draws)."""
        if str(self.corrs_device) != "cpu" or not any(valid):
            return None
        depth_a, depth_q = self._stacked_depth(batch["anchor"]), self._stacked_depth(batch["query"])
        if not (self._uniform_sizes(batch, "anchor", depth_a) and self._uniform_sizes(batch, "query", depth_q)):
            return None
        ma, mq, na, nq = self._masks_and_counts(results)
        fa, fq = outputs["featmap_a"], outputs["featmap_q"]
        B = fa.shape[0]
        roi_a, cnt_a = mask_to_roi(ma)
        roi_q, cnt_q = mask_to_roi(mq)
        n_a = [n if v else 0 for n, v in zip(cnt_a.tolist(), valid)]
        n_q = [n if v else 0 for n, v in zip(cnt_q.tolist(), valid)]
        idx, dist = match_nn(fa, fq, roi_a, roi_q, n_a, n_q)
        if self._host_dist is None or self._host_dist.shape != dist.shape:
            self._host_dist = torch.empty(dist.shape, dtype=torch.float32).pin_memory()
            self._host_rows = torch.empty(B, self.n_corrs, dtype=torch.int32).pin_memory()
        self._host_dist.copy_(dist, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()                  # sync 1: distances for the draws
        self._host_rows.copy_(self.draw_rows(self._host_dist, n_a, valid, batch.get("pair_index")))
        corrs, pa, pq, nv = select_lift_batched(self._host_rows, roi_a, roi_q, idx, depth_a, depth_q, batch["anchor"]["camera"],
                                                batch["query"]["camera"], self.featmap_size)
        counts = nv.tolist()                                                    # sync 2: points that survived the bounds test
        todo = [b for b in range(B) if counts[b] >= 0]
        poses = {}
        if todo:
            if self.solver != "pointdsc":
                raise RuntimeError(f"Solver {self.solver} not implemented")
            if len(todo) == B:
                src, tgt = pa, pq
            else:
                sel = torch.tensor(todo, device=self.device)
                src, tgt = pa.index_select(0, sel), pq.index_select(0, sel)
            T = pointdsc_poses(self.pointdsc_solver.to(self.device), src, tgt, counts=[counts[b] for b in todo]).cpu().to(torch.float32)
            poses = {b: T[i] for i, b in enumerate(todo)}
        return [corrs[b] if counts[b] >= 0 else None for b in range(B)], poses

    def test_step(self, batch: dict, batch_idx: int = 0, *, register_as_test: bool = True) -> List[dict]:
        """The reference's hot loop (pipeline.py:306-355) for one batch; returns (and accumulates in ``self.rows``)
        one record per pair: ids, ``pred_pose_rel`` (identity on failure, :335-350), ``pred_pose`` =
        ``pred_pose_rel @ anchor pose`` (:320), IoUs, status, and writes the CSV line when a file is open.

        ``test.pipelined``: the step is split into a FRONT (network pass + mask post-processing, enqueued on the current
        stream) and a BACK (matching -> draws -> selection / lifting -> PointDSC -> rows).  The call enqueues the front of ITS
        batch, then runs the back of the PREVIOUS batch on a second, high-priority stream -- its kernels and all of its host
        work (two host synchronisations, the CPU-generator draws, the row bookkeeping) execute while the GPU is busy with the
        network pass just enqueued -- and returns the previous batch's rows (``[]`` on the first call).  ``on_test_end`` /
        ``flush`` finish the last batch.  CSV lines, evaluator registrations and the draws on the CPU generator keep the
        reference's pair order, so a run produces the same files as the unpipelined loop; like the reference's own
        ``test_step`` (which returns ``None``) the results of a batch are only complete at ``on_test_end``."""
        front = self._front(batch, register_as_test)
        if not self.pipelined:
            return self._back(front)
        if self._tail_stream is None:
            self._tail_stream = torch.cuda.Stream(self.device, priority=-1)
        front["ready"] = torch.cuda.Event()
        front["ready"].record(torch.cuda.current_stream(self.device))
        pending, self._pending = self._pending, front
        return self._back(pending) if pending is not None else []

    def _front(self, batch: dict, register_as_test: bool = True) -> dict:
        outputs = self.forward(batch)
        results = self.mask_results(batch, outputs)
        return dict(batch=batch, outputs=outputs, results=results, register_as_test=register_as_test)

    def _back(self, front: dict) -> List[dict]:
        if front.get("ready") is None:
            return self._back_on_current_stream(front)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._tail_stream):
            self._tail_stream.wait_event(front["ready"])
            rows = self._back_on_current_stream(front)
            self._tail_stream.synchronize()      # everything that read this batch's tensors has run: they may be released
        # tensors created by the back on the tail stream and handed to the caller (rows[i]['corrs']) are safe to use on `main`
        main.wait_stream(self._tail_stream)
        return rows

    def _back_on_current_stream(self, front: dict) -> List[dict]:
        batch, outputs, results, register_as_test = front["batch"], front["outputs"], front["results"], front["register_as_test"]
        B = outputs["featmap_a"].shape[0]
        _, _, na, nq = self._masks_and_counts(results)
        valid = [(a > 0 and q > 0) for a, q in zip(na.tolist(), nq.tolist())]
        tail = self._batched_tail(batch, outputs, results, valid) if self.batched_tail else None
        if tail is not None:
            corrs, poses = tail
        else:
            corrs = self.select_correspondences(results, outputs, valid, batch.get("pair_index"))
            # lifting per pair, registration for all pairs with correspondences at once
            todo, pa, pq = [], [], []
            for b in range(B):
                if corrs[b] is not None:
                    a, q = self._lift(batch, corrs[b], b)
                    todo.append(b), pa.append(a), pq.append(q)
            poses = {}
            if todo:
                if self.solver != "pointdsc":
                    raise RuntimeError(f"Solver {self.solver} not implemented")
                T = pointdsc_poses(self.pointdsc_solver.to(self.device), pa, pq).cpu().to(torch.float32)
                poses = {b: T[i] for i, b in enumerate(todo)}
        iou_a = results["iou_a"].cpu().numpy() if results["iou_a"] is not None else np.full(B, np.nan, np.float32)
        iou_q = results["iou_q"].cpu().numpy() if results["iou_q"] is not None else np.full(B, np.nan, np.float32)
        rows = []
        for b in range(B):
            id_a, id_q = batch["anchor"]["instance_id"][b], batch["query"]["instance_id"][b]
            if b in poses:
                pred_pose, status = poses[b], "ok"
                pred_q = pred_pose @ batch["anchor"]["pose"][b].cpu().detach().to(torch.float32)
            else:
                pred_pose, status = torch.eye(4), ("no_corrs" if valid[b] else "invalid_mask")
                pred_q = None
            self.add_pred_pose(id_a, id_q, iou_a[b], iou_q[b], pred_pose.cpu().numpy())
            rows.append(dict(instance_id_a=id_a, instance_id_q=id_q, pred_pose_rel=pred_pose, pred_pose=pred_q, iou_a=float(iou_a[b]),
                             iou_q=float(iou_q[b]), status=status, corrs=corrs[b]))
        self.rows.extend(rows)
        if self.evaluator is not None:
            self._register(batch, rows, test=register_as_test)
        return rows

    def validation_step(self, batch: dict, batch_idx: int = 0) -> List[dict]:
        """The inference part of the reference's ``validation_step`` (pipeline.py:196-247): the same per-pair path as
        ``test_step``, registered through ``register_eval`` / ``register_valid_failure`` (no instance ids), no prediction CSV.
        The loss terms of that hook (``FeatureLoss``) are training code and are not computed: the records are returned instead
        of a loss."""
        self.flush()                                     # a validation batch is never deferred (no CSV, rows returned at once)
        pred_file, self.pred_file = self.pred_file, None
        pipelined, self.pipelined = self.pipelined, False
        try:
            return self.test_step(batch, batch_idx, register_as_test=False)
        finally:
            self.pred_file, self.pipelined = pred_file, pipelined

    @staticmethod
    def _eval_depth(batch: dict, i: int):
        """The query depth frame the evaluator's VSD term reads (``batch['query']['eval_depth'][i]``, pipeline.py:328), or None."""
        frames = batch["query"].get("eval_depth") if isinstance(batch.get("query"), dict) else None
        if frames is None:
            return None
        d = frames[i]
        return d.squeeze().cpu().numpy() if isinstance(d, Tensor) else np.asarray(d).squeeze()

    def _register(self, batch: dict, rows: List[dict], test: bool = True) -> None:
        """Evaluator bookkeeping of the reference's loop (pipeline.py:321-350; :207-245 for validation), in pair order;
        consecutive successful pairs go to the (batched, GPU) ``register_test`` / ``register_eval`` in one call."""
        def flush(run):
            if not run:
                return
            sel = torch.tensor(run)
            payload = {
                "iou_a": torch.tensor([rows[i]["iou_a"] for i in run]), "iou_q": torch.tensor([rows[i]["iou_q"] for i in run]),
                "gt_pose": batch["query"]["pose"].cpu()[sel], "pred_pose": torch.stack([rows[i]["pred_pose"] for i in run]),
                "pred_pose_rel": torch.stack([rows[i]["pred_pose_rel"] for i in run]), "cls_id": [batch["cls_id"][i] for i in run],
                "camera": [batch["query"]["camera"][i].cpu().numpy() for i in run], "depth": [self._eval_depth(batch, i) for i in run]}
            if test:
                payload["instance_id"] = [batch["instance_id"][i] for i in run]
                self.evaluator.register_test(payload)
            else:
                self.evaluator.register_eval(payload)

        run: List[int] = []
        for i, r in enumerate(rows):
            if r["status"] == "ok":
                run.append(i)
                continue
            flush(run)
            run = []
            failure = {"iou_a": torch.tensor([r["iou_a"]]), "iou_q": torch.tensor([r["iou_q"]]), "cls_id": [batch["cls_id"][i]],
                       "instance_id": [batch["instance_id"][i]]}
            if test:
                self.evaluator.register_test_failure(failure)
            else:
                self.evaluator.register_valid_failure(failure)
        flush(run)
