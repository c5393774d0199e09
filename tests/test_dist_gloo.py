"""CPU, world_size 2 and 4, gloo: the multi-GPU host logic (shard, all-gather of the result tables)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scene_util as su


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, os.path.join(su.ROOT, "diff-dope_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffdope import _dist

    lo, hi = _dist.shard_range(B)
    n = 3
    pose = torch.arange(n * B * 7, dtype=torch.float32).reshape(n, B, 7)[:, lo:hi].contiguous()
    loss = torch.arange(n * B * 3, dtype=torch.float32).reshape(n, B, 3)[:, lo:hi].contiguous() * 0.5
    final = torch.arange(B * 7, dtype=torch.float32).reshape(B, 7)[lo:hi].contiguous() + 100
    flat = torch.cat([pose.reshape(-1), loss.reshape(-1), final.reshape(-1)])  # a rank's flat result buffer (DiffDope._fused_enqueue)
    assert flat.numel() == _dist.flat_sizes(n, hi - lo, 3)[2]
    P, L, F = _dist.gather_hypotheses(B, n, 3, flat)
    lr = torch.full((B,), float(rank + 1))
    _dist.broadcast_from_rank0(lr)
    torch.save({"P": P, "L": L, "F": F, "lr": lr, "range": (lo, hi)}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def _run_world(tmp_path, world, B):
    port = 29500 + (os.getpid() + 17 * B + 101 * world) % 2000
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "r%d.pt" % r)) for r in range(world)]
    n = 3
    P = torch.arange(n * B * 7, dtype=torch.float32).reshape(n, B, 7)
    L = torch.arange(n * B * 3, dtype=torch.float32).reshape(n, B, 3) * 0.5
    F = torch.arange(B * 7, dtype=torch.float32).reshape(B, 7) + 100
    for r in res:
        assert torch.equal(r["P"], P) and torch.equal(r["L"], L) and torch.equal(r["F"], F)
        assert torch.all(r["lr"] == 1.0)  # rank 0's draw is the job's
    # the shards tile [0, B) in rank order; trailing ranks may be empty when B < world or B is ragged
    assert res[0]["range"][0] == 0 and res[-1]["range"][1] == B
    for a, b in zip(res[:-1], res[1:]):
        assert a["range"][1] == b["range"][0] and a["range"][0] <= a["range"][1]
    # every rank computes the same argmin from the gathered table
    assert len({int(r["L"][-1].mean(-1).argmin()) for r in res}) == 1
    return res


def test_gather_hypotheses_world2(tmp_path):
    for B in (8, 7):
        _run_world(tmp_path, 2, B)


def test_gather_hypotheses_world4_ragged_and_empty_shards(tmp_path):
    res = _run_world(tmp_path, 4, 5)  # shards of 2, 2, 1, 0 hypotheses
    assert [r["range"] for r in res] == [(0, 2), (2, 4), (4, 5), (5, 5)]
    res = _run_world(tmp_path, 4, 2)  # fewer hypotheses than ranks: two empty shards
    assert [r["range"][1] - r["range"][0] for r in res] == [1, 1, 0, 0]
