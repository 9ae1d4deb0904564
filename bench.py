#!/usr/bin/env python
"""bench.py — subgraphs/sec per train step (all 3 channels) of the B200-native SubGNN hot path.

    python bench.py --gpus N --steps K --warmup W [--workload ppi_bp] [--impl reference]

One "step" = forward + loss + backward + gradient clipping + Adam on one batch of synthetic subgraphs
(batch size from the reference's hyper-parameter file).  Default workload: the PPI-BP-shaped all-channel
configuration (BASELINE.json configs[2], "all channels on 1 B200"); weak scaling for N > 1 (every rank
steps its own batch, gradients all-reduced over NCCL).  Prints ONE JSON line on rank 0.

--impl reference times the CPU oracle port of the reference's per-step path (oracle/model.py; the reference
itself is pure Python and absent on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'subgraphs_per_sec_train_step_all_channels'
UNIT = 'subgraphs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default='ppi_bp')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch-size', type=int, default=0, help='override the per-GPU batch (default: reference hparams)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-extra', dest='extra', action='store_false', help='skip the other BASELINE configs / strong / saturating legs')
    ap.add_argument('--no-anomaly', dest='anomaly', action='store_false', help='skip the anomaly-detection-on CPU leg')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).
    The poller is started before the warm-up (nvidia-smi needs ~100 ms to come up) and every sample is stamped on
    arrival; only samples that fall inside [mark_begin, mark_end] windows (GPU under the bench load) are reported."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index, period_ms=50):
        self.gpu, self.rows, self.proc, self.period_ms, self.windows = gpu_index, [], None, period_ms, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(',')]))

    def mark_begin(self):
        self._t0 = time.perf_counter()

    def mark_end(self):
        self.windows.append((self._t0, time.perf_counter()))

    def in_window(self):
        # a sample printed at time t describes the preceding polling period
        return [r for t, r in self.rows if len(r) >= 8 and any(a + 1e-3 * self.period_ms <= t <= b + 1e-3 * self.period_ms for a, b in self.windows)]

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.in_window()
        num = lambda x: x.replace('.', '', 1).isdigit()
        sm = [float(r[1]) for r in rows if num(r[1])]
        mx = [float(r[2]) for r in rows if num(r[2])]
        pw = [float(r[3]) for r in rows if num(r[3])]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm), 'power_w': float(np.median(pw)) if pw else None}


# ----------------------------------------------------------------------------------------------------
def build_workload(name, device, batch_override=0):
    from subgnn_b200 import prepare as prep
    from subgnn_b200 import synth
    hp, g, subs, labs, emb = synth.make_workload(name, seed=42, device=device)
    if batch_override:
        hp['batch_size'] = batch_override
    t0 = time.time()
    prepared = prep.prepare(hp, g, subs, labs, emb, seed=0, splits=('train',), num_classes=synth.WORKLOADS[name]['n_classes'])
    import torch
    torch.cuda.synchronize()
    return hp, g, prepared, time.time() - t0


def batches_for(n_train, B, n_steps, rank, world, seed=0):
    """contiguous per-rank shards of a shuffled epoch stream (drop_last, SubGNN.py:1126-1127)."""
    rs = np.random.RandomState(seed)
    per_step = B * world
    out, perm, pos = [], rs.permutation(n_train), 0
    for _ in range(n_steps):
        if pos + per_step > n_train:
            perm, pos = rs.permutation(n_train), 0
        if per_step > n_train:                       # tiny data sets: sample with replacement across ranks
            chunk = rs.randint(n_train, size=per_step)
        else:
            chunk = perm[pos:pos + per_step]
            pos += per_step
        out.append(np.sort(chunk[rank * B:(rank + 1) * B]))
    return out


# algorithmic bytes / flops per launch of the entry points that can dominate a step (DESIGN.md "Roofline")
def algorithmic_work(eng, ctx):
    hp, t = eng.hp, ctx.tables
    D, L, B = hp['node_embed_size'], hp['n_layers'], ctx.B
    R = int(ctx.meta[0].item())
    idx = ctx.batch_idx.cpu().numpy()
    ccptr, nodeptr = t.sub_ccptr.cpu().numpy(), t.cc_nodeptr.cpu().numpy()
    n_nodes = int(sum(nodeptr[ccptr[i + 1]] - nodeptr[ccptr[i]] for i in idx))
    A = eng.A
    hid, h1, h2, K = eng.hid_dim, hp['linear_hidden_dim_1'], hp['linear_hidden_dim_2'], eng.num_classes
    useN = 1 if hp['use_neighborhood'] else 0
    mlp_w = 4 * (hid * h1 + h1 * h2 + h2 * K)
    fwd = (n_nodes * (4 + 4 * D) + R * 4 * D                                             # pooling: ids + rows, X0 write
           + useN * R * L * (A['ni'] + A['nb']) * (8 + 4 * D)                            # N gathers: id + sim + row
           + useN * R * L * 2 * (2 * 4 * D) + useN * L * 2 * (2 * D * D + D) * 4         # Nagg/Nh writes, weights (once)
           + R * L * (A['pi'] + A['pb'] + 2 * A['s']) * 4                                # P/S similarities
           + L * (B * A['pi'] + A['pb'] + 2 * A['s']) * 4                                # q
           + B * hid * 4 + mlp_w + B * (h1 + h2 + K) * 4 * 2)                            # Z, MLP weights (once), activations
    bwd = (B * hid * 4
           + useN * R * L * 2 * (3 * 4 * D) + useN * L * 2 * 2 * D * D * 4               # Nh read, dpre write, weights
           + useN * R * L * (A['ni'] + A['nb']) * (8 + 8 * D)                            # scatter: id + sim + row RMW
           + n_nodes * (4 + 8 * D)                                                        # pooling scatter RMW
           + R * L * (A['pi'] + A['pb'] + 2 * A['s']) * 4 + L * (B * A['pi'] + A['pb'] + 2 * A['s']) * 8)
    n_par = eng.arena.size
    # amounts are PER STEP (all launches of the entry point in one step together)
    work = {'subgnn_model_rows_fwd': ('hbm', fwd), 'subgnn_model_rows_bwd': ('hbm', bwd),
            'subgnn_adam_step': ('hbm', 28 * n_par), 'subgnn_grad_sumsq': ('hbm', 4 * n_par), 'subgnn_fill_zero': ('hbm', 4 * n_par)}
    if eng.lstm is not None:
        ls = eng.lstm
        M, H = ls.n_seq * ls.T, ls.H
        flops_in = sum(2 * M * 8 * H * (D if k == 0 else 2 * H) for k in range(ls.nl))
        # 'last' aggregator: the top layer's reverse direction takes one step only (SubGNN.py:83)
        top_rev = 0 if ls.sum_mode else 1
        rows_dir = [[M, M] for _ in range(ls.nl)]
        if top_rev:
            rows_dir[-1][1] = ls.n_seq
        flops_proj = sum(2 * (rows_dir[k][0] + rows_dir[k][1]) * 4 * H * (D if k == 0 else 2 * H) for k in range(ls.nl))
        flops_rec = sum(2 * (rows_dir[k][0] + rows_dir[k][1]) * 4 * H * H for k in range(ls.nl))
        # algorithmic bytes of the same launches (every operand read once, every result written once, fp32): the GEMMs have
        # K <= 2H <= 256, i.e. an arithmetic intensity of N K / (2 (N + K)) ~ 28-60 flop/B, far below the ridge of the part
        # (measured tensor peak / measured HBM peak ~ 210 flop/B): by the roofline they are HBM-bound, not tensor-bound
        rows_all = sum(rows_dir[k][0] + rows_dir[k][1] for k in range(ls.nl))
        b_proj = sum(4 * (M * (D if k == 0 else 2 * H) + (rows_dir[k][0] + rows_dir[k][1]) * 4 * H + 8 * H * (D if k == 0 else 2 * H)) for k in range(ls.nl))
        b_whh = 4 * (rows_all * (4 * H + H) + ls.nl * 8 * H * H)
        b_rec_f = 4 * rows_all * (4 * H + 4 * H + H + H) + 4 * ls.nl * 8 * H * H          # G in, gates out, c, h; W_hh once
        b_rec_b = 4 * rows_all * (4 * H + 4 * H + 2 * H + 2 * H) + 4 * ls.nl * 8 * H * H  # gates, dG out, c_t / c_{t-1}, h / dh
        # the TMA-fed grouped kernel carries every dense product of the LSTM layers: input projections, their input gradients
        # (layer 0: row scatter into dE, only with trainable embeddings) and input-weight gradients, and the recurrent-weight gradients;
        # bytes: every product reads its operands once and writes its result once (a grouped launch re-reads dG from L2)
        n_bi = ls.nl if eng.dE_ptr() else ls.nl - 1
        flops_bi = sum(2 * (rows_dir[k][0] + rows_dir[k][1]) * 4 * H * (D if k == 0 else 2 * H) for k in range(ls.nl) if (k > 0 or eng.dE_ptr()))
        b_bi = sum(4 * (M * (D if k == 0 else 2 * H) + (rows_dir[k][0] + rows_dir[k][1]) * 4 * H + 8 * H * (D if k == 0 else 2 * H))
                   for k in range(ls.nl) if (k > 0 or eng.dE_ptr()))
        work['subgnn_tc_gemm_group'] = ('gemm', 2 * flops_proj + flops_bi + flops_rec, 2 * b_proj + b_bi + b_whh)
        work['subgnn_tc_linear_fwd'] = ('gemm', flops_proj, b_proj)
        work['subgnn_tc_linear_bwd_weight'] = ('gemm', flops_proj + flops_rec, b_proj + b_whh)   # d W_ih and d W_hh
        work['subgnn_tc_linear_bwd_input'] = ('gemm', flops_proj, b_proj)
        work['subgnn_lstm_recur_fwd'] = ('fp32', flops_rec, b_rec_f)
        work['subgnn_lstm_recur_bwd'] = ('fp32', flops_rec, b_rec_b)
    return work, {'rows': R, 'component_nodes': n_nodes}


def measured_peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        p = json.loads(f.read_text())
        return {'hbm': p['hbm_gbs'], 'tensor': p['bf16_tflops'], 'tensor_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'src': 'measured'}
    return {'hbm': 6650.0, 'tensor': 1590.0, 'tensor_sustained': 1400.0, 'src': 'fallback'}


# entry point -> CUDA kernel(s) it launches, for the ncu traffic table (tools/ncu_traffic.py -> profiles/r01_traffic_<workload>.json)
ENTRY_KERNELS = {'subgnn_tc_gemm_group': ['tc_gemm_ws_kernel'], 'subgnn_tc_linear_bwd_weight': ['tc_linear_bwd_weight_kernel'], 'subgnn_tc_linear_fwd': ['tc_linear_fwd_kernel'],
                 'subgnn_tc_linear_bwd_input': ['tc_linear_bwd_input_kernel'], 'subgnn_lstm_recur_fwd': ['lstm_fwd_tile_kernel'],
                 'subgnn_lstm_recur_bwd': ['lstm_bwd_tile_kernel'], 'subgnn_model_rows_fwd': ['row_fwd_kernel'],
                 'subgnn_model_rows_bwd': ['row_bwd_kernel'], 'subgnn_adam_step': ['adam_kernel'], 'subgnn_grad_sumsq': ['sumsq_kernel'],
                 'subgnn_fill_zero': ['fill_zero_kernel']}


# source files that define each entry point's kernels (plus the shared device headers): the stamp checked by ncu_traffic
ENTRY_SOURCES = {'subgnn_tc_gemm_group': ['tcgemm_ws.cu'], 'subgnn_tc_linear_bwd_weight': ['tcgemm.cu'], 'subgnn_tc_linear_fwd': ['tcgemm.cu'],
                 'subgnn_tc_linear_bwd_input': ['tcgemm.cu'], 'subgnn_lstm_recur_fwd': ['lstm_reg.cu', 'lstm_reg.cuh'],
                 'subgnn_lstm_recur_bwd': ['lstm_reg.cu', 'lstm_reg.cuh'], 'subgnn_model_rows_fwd': ['model.cu'], 'subgnn_model_rows_bwd': ['model.cu'],
                 'subgnn_adam_step': ['optim.cu'], 'subgnn_grad_sumsq': ['optim.cu'], 'subgnn_fill_zero': ['optim.cu']}


def ncu_traffic(workload, entry):
    """mean DRAM bytes per launch of the entry point's kernel from the committed ncu --set full capture of this workload, or None.
    The table is stamped with the digest of every CUDA source it was captured on (tools/ncu_traffic.py); it is refused
    (traffic = null) when the files that define THIS entry point's kernels (ENTRY_SOURCES + common.cuh) have changed since."""
    f = ROOT / 'profiles' / ('r02_traffic_%s.json' % workload)
    if not f.exists() or entry not in ENTRY_KERNELS:
        return None, None
    tab = json.loads(f.read_text())
    stamps, now = tab.get('csrc_files') or {}, csrc_file_digests()
    for src in ENTRY_SOURCES.get(entry, sorted(now)) + ['common.cuh']:
        if src not in stamps or stamps[src] != now.get(src):
            return None, None
    ks = [tab['kernels'][k] for k in ENTRY_KERNELS[entry] if k in tab['kernels']]
    if not ks:
        return None, None
    n = sum(k['launches_per_step'] for k in ks)
    return sum(k['dram_bytes_per_launch'] * k['launches_per_step'] for k in ks) / n, 'profiles/%s (%s)' % (f.name, tab.get('source', ''))


# fp32 FFMA issue peak of the part (not in MEASURED_PEAKS.json): 148 SMs x 128 lanes x 2 flop x 1.965 GHz
FP32_FFMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


# ----------------------------------------------------------------------------------------------------
def cpu_baseline(hp, g, prepared, seconds, n_threads=None, anomaly=False):
    """The oracle port of the reference step on a bounded sample: 4 batches of the workload's train split,
    cycled for ~`seconds` of CPU work.  Returns (subgraphs/s incl. batch assembly, dict)."""
    import torch
    from oracle.model import OracleSubGNN
    from subgnn_b200 import prepare as prep
    if n_threads:
        torch.set_num_threads(n_threads)
    B = hp['batch_size']
    n_train = len(prepared['labels']['train'])
    n_sample = min(n_train, 4 * B)
    sample = prep.prepared_subset(prepared, g, 'train', np.arange(n_sample))
    torch.manual_seed(0)
    model = OracleSubGNN(hp, sample)
    opt = torch.optim.Adam(model.parameters(), lr=hp['learning_rate'])
    model.train()
    B = min(B, n_sample)
    order = [np.arange(i, i + B) for i in range(0, n_sample - B + 1, B)]
    n_done, t_used = 0, 0.0

    def one(idx):
        batch = model.make_batch('train', idx)                          # SubgraphDataset + _pad_collate
        loss, _ = model.training_step(batch)
        opt.zero_grad()
        loss.backward(retain_graph=True)
        torch.nn.utils.clip_grad_norm_(model.parameters(), hp['grad_clip'])
        opt.step()
        return float(loss.detach())

    with torch.autograd.set_detect_anomaly(anomaly):
        one(order[0])                                                   # warm-up
        t0 = time.perf_counter()
        i = 0
        while True:
            one(order[i % len(order)])
            n_done += 1
            i += 1
            t_used = time.perf_counter() - t0
            if t_used >= seconds or n_done >= 200:
                break
    value = n_done * B / t_used
    info = {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d steps of batch %d over the first %d train subgraphs of %s (oracle/model.py: reference op sequence incl. '
                      'SubgraphDataset/_pad_collate batch assembly, anomaly detection %s), %.1f s' %
                      (n_done, B, n_sample, 'the workload', 'on' if anomaly else 'off', t_used),
            'ms_per_step': 1e3 * t_used / n_done}
    return value, info


# ----------------------------------------------------------------------------------------------------
_RESULT_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line (rank 0).  Native libraries write there too (NCCL prints its version banner to fd 1
    whatever NCCL_DEBUG says under torchrun), so fd 1 is pointed at stderr for the life of the process and the result line is
    written to a private duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def config_for(workload, hp, g, n_train, world, B, use_graph=True):
    """the workload description both arms print (the driver compares the two dicts)."""
    return {'workload': (workload + '-shaped synthetic (SURVEY 8d), all channels N+P+S') if workload != 'cutratio' else 'cutratio-shaped synthetic, S only',
            'batch_per_gpu': B, 'global_batch': B * world, 'n_layers': hp['n_layers'], 'node_embed_size': hp['node_embed_size'],
            'graph_nodes': g.n_nodes, 'graph_edges': int(g.col.numel() // 2), 'train_subgraphs': n_train, 'parallelism': 'dp%d' % world,
            'l2': 'flushed between timed iterations (256 MB write)', 'cuda_graph': bool(use_graph)}


def csrc_digest():
    """sha256 over the CUDA sources: stamps the ncu traffic tables (a table captured on other kernels is refused)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / 'subgnn_b200' / 'csrc').glob('*.cu*')):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def csrc_file_digests():
    """{file name: sha256[:16]} of every CUDA source / header under csrc/"""
    import hashlib
    return {f.name: hashlib.sha256(f.read_bytes()).hexdigest()[:16] for f in sorted((ROOT / 'subgnn_b200' / 'csrc').glob('*.cu*'))}


class Job:
    """process-wide state shared by the measurement legs (rank / world / device / flush buffer / clock sampler)."""

    def __init__(self, args):
        import torch
        self.args = args
        self.rank = int(os.environ.get('RANK', 0))
        self.local_rank = int(os.environ.get('LOCAL_RANK', 0))
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        torch.cuda.set_device(self.local_rank)
        self.dev = 'cuda:%d' % self.local_rank
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device(self.dev))
            self.dist = dist
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)      # 256 MB > 126 MB L2
        self.sampler = ClockSampler(self.local_rank)
        self.sampler.start()

    def barrier(self):
        import torch
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        if self.dist is None:
            return float(x)
        tm = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
        return float(tm.item())


def measure(job, workload, K, W, batch_override=0, strong=False, detail=True, e2e_leg=True):
    """One workload on this job's ranks: W warm-up steps, K timed steps (CUDA events per step on the launching stream, L2 flushed
    in between, max over ranks), the e2e leg through SubGNN.training_step_fused with host batches, and — on rank 0, detail=True —
    the instrumented per-entry-point pass for the roofline.  strong: batch_override is the GLOBAL batch, split over the ranks."""
    import torch
    from subgnn_b200 import _abi
    from subgnn_b200.engine import Engine
    from subgnn_b200.SubGNN import SubGNN
    args, world, rank, dev = job.args, job.world, job.rank, job.dev
    B_over = batch_override
    if strong and batch_override:
        assert batch_override % world == 0, 'strong scaling: the global batch must divide over the ranks'
        B_over = batch_override // world
    hp, g, prepared, prep_s = build_workload(workload, dev, B_over)
    B = hp['batch_size']
    eng = Engine(hp, prepared, device=dev, graph=g, seed=1234, world_size=world)
    eng.init_parameters(seed=7)
    n_train = len(prepared['labels']['train'])
    use_graph = not args.no_graph
    batches = batches_for(n_train, B, W + K + 2, rank, world)
    flush, sampler = job.flush, job.sampler
    # ---- warm-up (includes graph capture) ----
    for i in range(W):
        eng.train_step(batches[i], use_graph=use_graph)
    if use_graph:
        eng.train_step(batches[W], use_graph=True)
    job.barrier()
    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    n0 = _abi.lib.subgnn_launch_count()
    job.barrier()
    sampler.mark_begin()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()
        idx = batches[W + 1 + i]
        evs[i][0].record()
        eng.train_step(idx, use_graph=use_graph)
        evs[i][1].record()
    job.barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    launches = int(_abi.lib.subgnn_launch_count() - n0)
    if use_graph:
        launches = int(getattr(eng, 'launches_per_step', 0)) * K
    total_ms = job.max_over_ranks(float(step_ms.sum()))
    value = world * B * K / (total_ms * 1e-3)
    final_loss = float(eng.context('train', B, True).loss.item())
    out = {'value': value, 'ms_per_step': total_ms / K, 'gpu_launches': launches, 'final_loss': final_loss, 'wall_s_timed_region': wall,
           'config': config_for(workload, hp, g, n_train, world, B, use_graph), 'prepare_data_s': round(prep_s, 2),
           'total_ms': total_ms}

    # ---- e2e: the user-facing call with host buffers; H2D of the step inputs + D2H of the loss inside the timed region ----
    e2e_s = 0.0
    if e2e_leg:
        model = SubGNN.from_engine(eng)
        host_batches = [{'subgraph_idx': torch.from_numpy(b.astype(np.int64)).view(-1, 1).pin_memory()} for b in batches[W + 1:W + 1 + K]]
        model.training_step_fused(host_batches[0], use_graph=use_graph)
        job.barrier()
        sampler.mark_begin()
        t0 = time.perf_counter()
        for hb in host_batches:
            res = model.training_step_fused(hb, use_graph=use_graph, sync_loss=True)
            _ = float(res['loss'])                                        # device -> host read of the step result
        job.barrier()
        e2e_s = job.max_over_ranks(time.perf_counter() - t0)
        sampler.mark_end()
        out['e2e'] = {'value': world * B * K / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': 4 * B, 'd2h_bytes_per_step': 4,
                      'note': 'SubGNN.training_step_fused(host batch dict): pinned H2D of the subgraph indices, fused step graph ending in a D2H '
                              'copy node of the loss, stream sync + host read every step (wall clock, so no L2 flush '
                              'inside this loop: a flush would be timed as work); all tables are device-resident after prepare_data (the reference re-uploads the dense similarity '
                              'slab every step)'}
    out['loaded_s'] = (total_ms * 1e-3) + e2e_s

    if detail and rank == 0:
        # ---- instrumented pass: per-entry-point device time with CUDA events (eager launches, same batches, no exchange) ----
        prof = {}

        @contextlib.contextmanager
        def hook(name):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield
            b.record()
            prof.setdefault(name, []).append((a, b))

        n_prof = min(K, 10)
        ws, dp = eng.world_size, eng.dp                  # rank 0 alone from here on: no exchange (its barriers would wait for the peers)
        for i in range(n_prof):
            torch.cuda.synchronize()
            # head start for the host: ~1 ms of L2-flushing fills are queued first, so every launch of the eager step is already in
            # the stream when the GPU reaches it and the event pairs measure device time, not the host's launch latency
            for _ in range(24):
                flush.zero_()
            _abi._profile_hook = hook
            eng.world_size, eng.dp = 1, None
            try:
                eng.train_step(batches[W + 1 + i], use_graph=False)
            finally:
                eng.world_size, eng.dp = ws, dp
                _abi._profile_hook = None
        torch.cuda.synchronize()
        per_entry = {k: (sum(a.elapsed_time(b) for a, b in v) / n_prof, len(v) // n_prof) for k, v in prof.items()}
        ctx = eng.context('train', B, True)
        work, stats = algorithmic_work(eng, ctx)
        out['rows_last_batch'] = stats['rows']
        out['breakdown_ms'] = {k: {'ms_per_step': round(v[0], 4), 'calls_per_step': v[1]} for k, v in sorted(per_entry.items(), key=lambda kv: -kv[1][0])}
        out['roofline'] = roofline_for(workload, per_entry, work)
    out['_eng'] = (eng, hp, g, prepared, batches)
    return out


def roofline_for(workload, per_entry, work):
    peaks = measured_peaks()

    def roof(name):
        if name not in work or name not in per_entry:
            return None
        ms, calls = per_entry[name]
        bound, amount = work[name][0], work[name][1]
        per_launch_s = ms * 1e-3 / calls
        amount_launch = amount / calls
        if bound == 'hbm':
            ach = amount_launch / per_launch_s / 1e9
            r = {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': ach / peaks['hbm'],
                 'traffic': None, 'algorithmic_bytes_per_launch': amount_launch, 'launches_per_step': calls,
                 'us_per_launch': per_launch_s * 1e6, 'peak_source': peaks['src']}
        else:
            # GEMM-shaped entries: the roofline bound follows from the arithmetic intensity against the ridge of the measured peaks
            flops_launch, bytes_launch = amount_launch, work[name][2] / calls
            ai, ridge = flops_launch / bytes_launch, peaks['tensor_sustained'] * 1e12 / (peaks['hbm'] * 1e9)
            tf, gbs = flops_launch / per_launch_s / 1e12, bytes_launch / per_launch_s / 1e9
            r = {'kernel': name, 'launches_per_step': calls, 'us_per_launch': per_launch_s * 1e6, 'peak_source': peaks['src'],
                 'algorithmic_flops_per_launch': flops_launch, 'algorithmic_bytes_per_launch': bytes_launch,
                 'arithmetic_intensity_flop_per_byte': ai, 'ridge_flop_per_byte': ridge, 'traffic': None,
                 'achieved_tflops': tf, 'frac_of_tensor_peak': tf / peaks['tensor_sustained'], 'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / peaks['hbm']}
            if bound == 'fp32':
                r.update({'fp32_ffma_peak_tflops': FP32_FFMA_TFLOPS, 'frac_of_fp32_ffma_peak': tf / FP32_FFMA_TFLOPS,
                          'note': 'T dependent steps of an (n_seq x H)(H x 4H) fp32 FFMA product + 5H activations per sequence; latency bound: the '
                                  'nominal fp32 FFMA issue rate (148 SMs x 128 lanes x 2 x 1.965 GHz) is reported beside the two contract peaks'})
            else:
                r['note'] = ('tcgen05 kind::tf32, 3xTF32 error compensation (3 MMAs per algorithmic product), accumulator in TMEM; TMA-fed, warp-specialised, '
                             'grouped: all dense products of the LSTM layers (projections, input / weight gradients), launches of 1-5 products')
            if ai < ridge:
                r.update({'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm']})
            else:
                r.update({'bound': 'tensor', 'achieved': tf, 'peak': peaks['tensor_sustained'], 'unit': 'TFLOP/s', 'frac': tf / peaks['tensor_sustained']})
        r['traffic'], src = ncu_traffic(workload, name)
        if src:
            r['traffic_source'] = src + ': dram__bytes_read.sum + dram__bytes_write.sum per launch, cold caches (upper bound for the in-graph launch)'
        return r

    ranked = sorted(per_entry, key=lambda k: -per_entry[k][0])
    roofline = next((r for r in (roof(k) for k in ranked) if r), None)
    if roofline is not None:
        roofline['others'] = [r for r in (roof(k) for k in ('subgnn_model_rows_fwd', 'subgnn_model_rows_bwd', 'subgnn_tc_linear_fwd',
                                                             'subgnn_tc_linear_bwd_input', 'subgnn_tc_linear_bwd_weight', 'subgnn_lstm_recur_fwd',
                                                             'subgnn_lstm_recur_bwd', 'subgnn_adam_step')
                                          if k in per_entry and k != roofline['kernel']) if r]
    return roofline


def release(m):
    """drop a leg's engine / tables before the next workload is built."""
    import gc
    import torch
    m.pop('_eng', None)
    gc.collect()
    torch.cuda.empty_cache()


EXTRA_WORKLOADS = ('density', 'cutratio', 'hpo_metab', 'em_user')


def reference_arm(args):
    """--impl reference: the CPU oracle port of the reference step (oracle/model.py) on rank 0's host cores, bounded sample of the
    SAME workload and batch size; prints the same `config` dict as the b200 arm.  The workload tensors are prepared with the repo's
    CUDA setup kernels (walks, DTW, hop table) — the timed loop is CPU only."""
    import torch
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if rank != 0:
        return 0
    dev = 'cuda:0' if torch.cuda.is_available() else None
    if dev is None:
        emit({'impl': 'reference', 'unavailable': 'workload preparation needs the CUDA setup kernels; no GPU visible'})
        return 0
    torch.cuda.set_device(0)
    hp, g, prepared, _ = build_workload(args.workload, dev, args.batch_size)
    n_train = len(prepared['labels']['train'])
    secs = min(120.0, max(5.0, 1.0 * args.steps)) if args.steps != 200 else 20.0
    value, info = cpu_baseline(hp, g, prepared, secs, n_threads=os.cpu_count())
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': info['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': config_for(args.workload, hp, g, n_train, world, hp['batch_size'], not args.no_graph),
            'cpu_baseline': info, 'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    if args.anomaly:
        _, info_a = cpu_baseline(hp, g, prepared, min(secs, 10.0), n_threads=os.cpu_count(), anomaly=True)
        line['cpu_baseline_anomaly_on'] = info_a
    emit(line)
    return 0


def main():
    args = parse()
    _claim_stdout()
    import torch

    if args.impl == 'reference':
        return reference_arm(args)

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (there is no CPU fallback)'
    job = Job(args)
    world, rank = job.world, job.rank
    K, W = args.steps, max(args.warmup, 3)

    main_leg = measure(job, args.workload, K, W, args.batch_size)
    eng, hp, g, prepared, batches = main_leg['_eng']
    B = hp['batch_size']
    use_graph = not args.no_graph

    # the timed region of a launch-bound step can be shorter than nvidia-smi's polling period: keep the GPU under the
    # same load (same step, same flush; untimed) until the poller has at least 5 samples under load
    probe_steps = 0
    job.sampler.mark_begin()
    loaded_s = main_leg['loaded_s']                         # identical on every rank (max-reduced): same number of extra rounds
    rounds = 0 if loaded_s >= 0.8 else min(400, int(np.ceil((0.8 - loaded_s) / max(main_leg['total_ms'] * 1e-3, 1e-4))))
    for _ in range(rounds):
        for i in range(K):
            job.flush.zero_()
            eng.train_step(batches[W + 1 + i], use_graph=use_graph)
        probe_steps += K
    job.barrier()
    job.sampler.mark_end()

    cpu = cpu_anom = None
    if world == 1 and not args.no_cpu_baseline:
        _, cpu = cpu_baseline(hp, g, prepared, args.cpu_seconds)
        if args.anomaly:
            # the reference ships with torch.autograd.set_detect_anomaly(True) (train_config.py:206): reported beside anomaly-off
            _, cpu_anom = cpu_baseline(hp, g, prepared, min(args.cpu_seconds, 8.0), anomaly=True)
    release(main_leg)
    del eng, prepared, batches

    # ---- the other BASELINE.json configurations, strong scaling and the saturating batch (same harness, fewer extras) ----
    configs, strong, saturating = {}, None, None
    if args.extra and args.workload == 'ppi_bp' and not args.batch_size:
        Ke = max(10, min(K, 50))
        for name in EXTRA_WORKLOADS:
            m = measure(job, name, Ke, W, detail=True)
            release(m)
            configs[name] = {'value': m['value'], 'unit': UNIT, 'ms_per_step': m['ms_per_step'], 'steps': Ke, 'e2e': m['e2e']['value'],
                             'batch_per_gpu': m['config']['batch_per_gpu'], 'config': m['config'], 'prepare_data_s': m['prepare_data_s'],
                             'gpu_launches_per_step': m['gpu_launches'] // Ke}
            if m.get('roofline'):
                r = m['roofline']
                configs[name]['roofline'] = {k: r.get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'us_per_launch')}
        if world > 1 and B % world == 0:
            # strong scaling at the reference's global batch: the batch of B subgraphs is split over the ranks
            m = measure(job, args.workload, Ke, W, batch_override=B, strong=True, detail=False)
            release(m)
            strong = {'global_batch': B, 'batch_per_gpu': B // world, 'value': m['value'], 'ms_per_step': m['ms_per_step'], 'e2e': m['e2e']['value'],
                      'unit': UNIT, 'scaling': 'strong'}
        if world == 1:
            # the whole train split as one batch: the LSTM chain is batch independent, so this is the throughput ceiling of the step
            n_train = main_leg['config']['train_subgraphs']
            m = measure(job, args.workload, Ke, W, batch_override=n_train, detail=True)
            release(m)
            saturating = {'batch': n_train, 'value': m['value'], 'ms_per_step': m['ms_per_step'], 'e2e': m['e2e']['value'], 'unit': UNIT}
            if m.get('roofline'):
                saturating['roofline'] = {k: m['roofline'].get(k) for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'us_per_launch')}
                saturating['breakdown_ms'] = dict(list(m['breakdown_ms'].items())[:6])

    clocks = job.sampler.stop()
    clocks['windows'] = 'timed regions + e2e regions of every leg' + (' + %d further untimed steps of the same loop (timed region shorter than the polling period)' % probe_steps if probe_steps else '')

    if rank == 0:
        line = {'metric': METRIC, 'value': main_leg['value'], 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': main_leg['ms_per_step'],
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': main_leg['config'], 'roofline': main_leg.get('roofline'), 'cpu_baseline': cpu, 'e2e': main_leg['e2e'],
                'gpu_launches': main_leg['gpu_launches'], 'clocks': clocks, 'breakdown_ms': main_leg.get('breakdown_ms'),
                'final_loss': main_leg['final_loss'], 'wall_s_timed_region': main_leg['wall_s_timed_region'],
                'workload_stats': {'rows_last_batch': main_leg.get('rows_last_batch'), 'prepare_data_s': main_leg['prepare_data_s']}}
        if cpu_anom is not None:
            line['cpu_baseline_anomaly_on'] = cpu_anom
        if configs:
            line['configs'] = configs
        if strong:
            line['strong'] = strong
        if saturating:
            line['saturating'] = saturating
        emit(line)
    if job.dist is not None:
        job.dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
