"""Data-parallel parity on real GPUs (needs >= 2 devices; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`):
2 ranks x B subgraphs == 1 rank x 2B subgraphs on the parameters after 5 steps, for the single-graph step (NCCL all-reduce captured
inside the step graph) and the split form.  The host-side exchange logic is covered on CPU by tests/test_dp_gloo.py."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_ranks_equal_one_rank_with_twice_the_batch():
    import os
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), str(ROOT / 'tests' / 'dp_equiv_worker.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert 'DP-EQUIV-OK world=2' in r.stdout
