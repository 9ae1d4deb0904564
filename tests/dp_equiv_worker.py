"""Worker of tests/test_gpu_dp.py (launched with torch.distributed.run, one process per GPU): data-parallel equivalence of the
REAL engine — `world` ranks stepping B subgraphs each must end with the parameters of ONE rank stepping world * B subgraphs
(dropout 0): SubGNN.py:317-348 training_step + :1156-1164 Adam + Lightning's clip, gradients averaged by one NCCL all-reduce.
Checks every form of the data-parallel step: gradient exchange + sharded Adam over NVLink peer memory (csrc/dp.cu, the default),
NCCL all-reduce captured inside the single step graph, and two graphs around an eager all-reduce."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


MODES = ('nvlink_fused', 'single_graph', 'split_graph')


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    dist.init_process_group('nccl', device_id=torch.device(dev))
    from subgnn_b200 import prepare as prep
    from subgnn_b200 import synth
    from subgnn_b200.engine import Engine
    hp, g, subs, labs, emb = synth.make_workload('tiny', seed=42, device=dev, n_sub=240)
    hp = dict(hp, lin_dropout=0.0, lstm_dropout=0.0, batch_size=8)
    prepared = prep.prepare(hp, g, subs, labs, emb, seed=0, splits=('train',), num_classes=3)
    n = len(prepared['labels']['train'])
    B, steps = hp['batch_size'], 5
    rs = np.random.RandomState(3)
    global_batches = [np.sort(rs.choice(n, size=B * world, replace=False)) for _ in range(steps)]
    results = {}
    for mode in MODES:
        os.environ['SUBGNN_DP_SPLIT_GRAPH'] = '1' if mode == 'split_graph' else '0'
        os.environ['SUBGNN_DP_FUSED'] = '1' if mode == 'nvlink_fused' else '0'
        eng = Engine(hp, prepared, device=dev, graph=g, seed=5, world_size=world)
        eng.init_parameters(3)
        if mode == 'nvlink_fused':
            assert eng.dp is not None, 'symmetric memory must be available on an NVLink box (the NCCL path is a fallback, not the product)'
        losses = []
        for gb in global_batches:
            loss = eng.train_step(gb[rank * B:(rank + 1) * B], use_graph=True)     # eager warm-up, capture, then replays
            losses.append(float(loss.item()))
        torch.cuda.synchronize()
        assert len(eng.context('train', B, True).graph) == (2 if mode == 'split_graph' else 1)
        if mode == 'nvlink_fused':                      # sharded optimizer state reassembles to the single-rank moments (checked below)
            results['osd'] = eng.optimizer_state_dict()
        results[mode] = ({k: v.clone() for k, v in eng.arena.state_dict().items()}, losses)
        # every rank holds the same parameters after the exchange
        flat = eng.arena.params.clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for other in gathered:
            assert torch.equal(other, flat), 'ranks diverged (%s)' % mode
        mean_loss = torch.tensor(losses, device=dev, dtype=torch.float64)
        dist.all_reduce(mean_loss)
        results[mode + '_loss'] = (mean_loss / world).cpu().numpy()
    ok = True
    if rank == 0:
        single = Engine(dict(hp, batch_size=B * world), prepared, device=dev, graph=g, seed=5, world_size=1)
        single.init_parameters(3)
        ref_losses = [float(single.train_step(gb, use_graph=True).item()) for gb in global_batches]
        torch.cuda.synchronize()
        ref = single.arena.state_dict()
        for mode in MODES:
            got, _ = results[mode]
            np.testing.assert_allclose(results[mode + '_loss'], ref_losses, rtol=1e-4, err_msg=mode + ' loss')
            for k in ref:
                np.testing.assert_allclose(got[k].cpu().numpy(), ref[k].cpu().numpy(), rtol=1e-4, atol=1e-5, err_msg='%s %s' % (mode, k))
        osd, ref_osd = results['osd'], single.optimizer_state_dict()
        for i in ref_osd['state']:
            np.testing.assert_allclose(osd['state'][i]['exp_avg'].numpy(), ref_osd['state'][i]['exp_avg'].numpy(), rtol=1e-4, atol=1e-7, err_msg='exp_avg %d' % i)
        print('DP-EQUIV-OK world=%d' % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
