"""Multi-GPU plumbing of the inference path: (anchor, query) pairs are independent, so the
pair list is split across ranks with NO data-path collective; the only collective is one ``all_gather`` of fixed-size
result rows at the end (NCCL over NVLink on GPUs, gloo in the CPU tests), after which rank 0 restores the pair order
and writes the single CSV the reference's offline scorer reads (scripts/evaluation/compute_metrics.py:14-49).

The reference itself has no gather (pipeline.py:358-370 runs per process); without one, a multi-GPU ``run_test.py``
would produce one partial CSV per rank.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

ROW_FLOATS = 16  # pair_index, status, iou_a, iou_q, 12 pose entries (rows 0..2 of pred_pose_rel)
STATUS = {"ok": 0, "no_corrs": 1, "invalid_mask": 2}
STATUS_NAMES = {v: k for k, v in STATUS.items()}


def shard_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous, un-padded split of ``range(n_pairs)`` (a DistributedSampler would duplicate pairs to equalise
    the shards and corrupt the means).  Shard sizes differ by at most one."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def encode_rows(pair_indices: Sequence[int], rows: Sequence[dict]) -> Tensor:
    """Per-pair records of ``FPM_Pipeline.test_step`` -> float64 ``[n, 16]`` wire rows."""
    out = torch.zeros(len(rows), ROW_FLOATS, dtype=torch.float64)
    for i, (pi, r) in enumerate(zip(pair_indices, rows)):
        out[i, 0], out[i, 1] = float(pi), float(STATUS[r["status"]])
        out[i, 2], out[i, 3] = float(r["iou_a"]), float(r["iou_q"])
        out[i, 4:] = r["pred_pose_rel"][:3, :].reshape(12).double()
    return out


def decode_rows(t: Tensor) -> List[dict]:
    res = []
    for row in t:
        pose = torch.eye(4)
        pose[:3, :] = row[4:].reshape(3, 4).float()
        res.append(dict(pair_index=int(row[0]), status=STATUS_NAMES[int(row[1])], iou_a=float(row[2]), iou_q=float(row[3]),
                        pred_pose_rel=pose))
    return res


def gather_rows(local: Tensor, n_pairs: int, group=None) -> Tensor:
    """All ranks contribute their ``[n_local,16]`` rows; every rank gets the ``[n_pairs,16]`` table sorted by pair
    index.  One ``all_gather`` of equally sized (zero-padded) blocks: shards differ by at most one row."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[torch.argsort(local[:, 0])]
    cap = (n_pairs + world - 1) // world
    block = torch.full((cap, ROW_FLOATS), -1.0, dtype=torch.float64, device=local.device)
    block[:local.shape[0]] = local
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    table = torch.cat(blocks)
    table = table[table[:, 0] >= 0]
    if table.shape[0] != n_pairs:
        raise RuntimeError(f"gather_rows: {table.shape[0]} rows gathered for {n_pairs} pairs")
    return table[torch.argsort(table[:, 0])]
