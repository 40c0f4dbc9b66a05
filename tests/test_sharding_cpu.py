"""CPU, world_size 2, gloo: the multi-GPU host logic (pair sharding without padding, the single final all_gather of
result rows, pair-order restoration).  The GPU path uses the same functions over NCCL (bench.py under torchrun)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oryon_b200 import sharding


def test_shard_pairs_partition():
    for n in (0, 1, 7, 2000, 2001):
        for world in (1, 2, 3, 8):
            shards = [sharding.shard_pairs(n, r, world) for r in range(world)]
            flat = [i for s in shards for i in s]
            assert flat == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def _fake_row(i):
    g = torch.Generator().manual_seed(i)
    pose = torch.eye(4)
    pose[:3, :] = torch.randn(3, 4, generator=g)
    status = ("ok", "no_corrs", "invalid_mask")[i % 3]
    return dict(status=status, iou_a=float(i) / 10, iou_q=float("nan") if i % 5 == 0 else 0.5, pred_pose_rel=pose)


def _worker(rank, world, n_pairs, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_pairs(n_pairs, rank, world)
    local = sharding.encode_rows(list(mine), [_fake_row(i) for i in mine])
    table = sharding.gather_rows(local, n_pairs)
    ret[rank] = table
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [7, 10])
def test_gather_rows_world2_gloo(n_pairs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() + n_pairs) % 2000
    mp.spawn(_worker, args=(world, n_pairs, port, ret), nprocs=world, join=True)
    assert torch.equal(torch.nan_to_num(ret[0], nan=-7.0), torch.nan_to_num(ret[1], nan=-7.0))
    rows = sharding.decode_rows(ret[0])
    assert [r["pair_index"] for r in rows] == list(range(n_pairs))
    for i, r in enumerate(rows):
        ref = _fake_row(i)
        assert r["status"] == ref["status"] and torch.allclose(r["pred_pose_rel"], ref["pred_pose_rel"])
        assert (r["iou_q"] != r["iou_q"]) == (i % 5 == 0)
