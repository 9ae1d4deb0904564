#!/usr/bin/env python
"""Device time of the data-parallel exchange (torchrun, one process per GPU): reduce-scatter + sharded Adam + all-gather over NVLink peer
memory with the barriers inside the kernels (per-peer and NVLS multicast forms) against NCCL all-reduce + the two optimizer kernels
for the same arena.  Every piece is replayed from a CUDA graph of `reps` repetitions.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_bench.py [arena_floats]"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist.init_process_group('nccl', device_id=torch.device(dev))
    from subgnn_b200._abi import call, ptr
    from subgnn_b200.engine import DpExchange
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1583776            # PPI-BP arena
    n = (n + 3) // 4 * 4

    class Arena:
        pass

    alloc = DpExchange.allocator(dev)
    a = Arena()
    a.size = n
    a.params, a.grads = alloc(n).zero_(), alloc(n).normal_()
    a.m, a.v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step_dev = torch.ones(1, dtype=torch.int32, device=dev)
    sumsq = torch.zeros(1, device=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
    res = {}
    for mc in (0, 1):
        for bar in ('torch', 'kernel'):
            os.environ['SUBGNN_DP_MULTICAST'] = str(mc)
            os.environ['SUBGNN_DP_BARRIER'] = bar
            dp = DpExchange(a, world, rank, dev)
            res['whole exchange (multicast %s, %s barriers)' % ('on' if dp.mc_g else 'off', 'signal-pad kernel' if bar == 'torch' else 'in-kernel flag')] = \
                timeit(lambda st: dp.step(1e-3, step_dev, 0.2, st))
            if mc == 0 and bar == 'torch':
                res['  one signal-pad barrier'] = timeit(lambda st: dp.h_g.barrier(channel=0))
                res['  reduce_scatter kernel alone (unsynchronised)'] = timeit(lambda st: call('subgnn_dp_reduce_scatter', dp.pg, dp.ps, dp.pf, dp.mc_g, None, world, rank, n,
                                                                                                dp.shard, ptr(dp.gsum), st))
                res['  adam_allgather kernel alone (unsynchronised)'] = timeit(lambda st: call('subgnn_dp_adam_allgather', dp.pp, dp.pf, dp.mc_p, None, world, rank, n, dp.shard,
                                                                                                ptr(dp.gsum), ptr(a.m), ptr(a.v), 1e-3, 0.9, 0.999, 1e-8, ptr(step_dev),
                                                                                                ptr(dp.slots), 0.2, 1.0 / world, st))
    nccl = {
        'nccl all_reduce': lambda st: dist.all_reduce(a.grads),
        'sumsq + adam (local)': lambda st: (call('subgnn_grad_sumsq', ptr(a.grads), n, ptr(sumsq), st),
                                             call('subgnn_adam_step', ptr(a.params), ptr(a.grads), ptr(a.m), ptr(a.v), n, 1e-3, 0.9, 0.999, 1e-8, ptr(step_dev),
                                                  ptr(sumsq), 0.2, 1.0 / world, st)),
    }
    for name, fn in nccl.items():
        res[name] = timeit(fn)
    if rank == 0:
        print('world %d, arena %d floats (%.1f MB), shard %.2f MB' % (world, n, 4e-6 * n, 4e-6 * n / world))
        for k, v in res.items():
            print('  %-72s %7.1f us' % (k, v))
    dist.barrier()
    dist.destroy_process_group()


def timeit(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(s.cuda_stream)
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn(s.cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            dist.barrier()
            a.record(s)
            g.replay()
            b.record(s)
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps * 1e3)
    t = torch.tensor([best], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if __name__ == '__main__':
    main()
