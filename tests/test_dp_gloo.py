"""Data-parallel host logic on CPU: world_size-2 gloo process group.  Checks the exchange step of the path
(gradient arena all-reduce + 1/world averaging) and the batch sharding, without any GPU."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import bench
    from subgnn_b200.engine import Engine
    # the engine's exchange step on a stand-in arena (the real arena lives on the GPU)
    eng = Engine.__new__(Engine)
    eng.world_size = world
    eng.arena = type('A', (), {})()
    eng.arena.grads = torch.arange(10, dtype=torch.float32) * (rank + 1)
    eng.allreduce_grads()
    want = torch.arange(10, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(eng.arena.grads, want)
    # every rank derives the same permutation and takes its own contiguous shard
    mine = bench.batches_for(64, 4, 3, rank, world, seed=1)
    gathered = [None] * world
    dist.all_gather_object(gathered, [b.tolist() for b in mine])
    disjoint = all(len(set(sum((g[i] for g in gathered), []))) == 4 * world for i in range(3))
    out[rank] = ok and disjoint
    dist.destroy_process_group()


def test_gradient_allreduce_and_sharding_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
