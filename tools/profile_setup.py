#!/usr/bin/env python
"""prepare_data (setup-phase kernels) of one bench workload inside the cudaProfiler range, for

    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:'walk_|sample_rows|border_|hop_table|sp_min|degree_seq|dtw_batch' -o R python tools/profile_setup.py --workload ppi_bp
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='ppi_bp')
    a = ap.parse_args()
    from subgnn_b200 import ops, prepare, synth
    torch.cuda.set_device(0)
    hp, g, subs, labs, emb = synth.make_workload(a.workload, seed=42, device='cuda:0')
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    p = prepare.prepare(hp, g, subs, labs, emb, seed=0, splits=('train',), num_classes=synth.WORKLOADS[a.workload]['n_classes'])
    # the dense similarity slab of the first components (the engine resolves similarities by gather; the reference stores this slab)
    prepare.dense_np_sim(g, p['cc_ids']['train'][:64])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
