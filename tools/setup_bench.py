#!/usr/bin/env python
"""Secondary metrics of SURVEY 8d: throughput of the setup-phase kernels (anchor-patch sampling, similarity gamma) on the
benchmark shapes, each against the roofline that bounds it, with the CPU oracle timed beside it on a bounded sample.

    python tools/setup_bench.py [--workloads ppi_bp hpo_metab em_user] [--cpu-seconds 5] > profiles/r01_setup_bench.json

Per kernel: units/s (walk steps, draws, (component, patch) pairs, hop-table rows ...), algorithmic bytes per unit,
achieved GB/s and its fraction of the measured HBM peak.  Times are CUDA events over `reps` launches after a warm-up, inputs
resident in HBM; between repetitions a 256 MB buffer is written to flush L2.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    return (json.loads(f.read_text())['hbm_gbs'], 'measured') if f.exists() else (6650.0, 'fallback')


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), out


def run(name, cpu_seconds, reps=5):
    from oracle import gamma as og
    from oracle import walks as ow
    from subgnn_b200 import ops, prepare, synth
    from subgnn_b200.graph import ragged_from_padded
    dev = 'cuda'
    hbm, src = peaks()
    hp, g, subs, labs, emb = synth.make_workload(name, seed=42, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    L, W, T, Ls, beta = hp['n_layers'], hp['n_triangular_walks'], hp['random_walk_len'], hp['sample_walk_len'], hp['rw_beta']
    P_tot = hp['max_sim_epochs'] * hp['n_anchor_patches_structure'] * L
    deg = (g.rowptr[1:] - g.rowptr[:-1]).float()
    res = {'workload': name, 'graph_nodes': g.n_nodes, 'graph_edges': int(g.col.numel() // 2), 'hbm_peak_gbs': hbm, 'peak_source': src, 'kernels': {}}

    def rec(key, seconds, units, unit_name, bytes_total, note='', cpu=None):
        e = {'seconds': seconds, 'units': units, 'unit': unit_name, 'units_per_s': units / seconds, 'algorithmic_bytes': bytes_total,
             'achieved_gbs': bytes_total / seconds / 1e9, 'frac_of_hbm_peak': bytes_total / seconds / 1e9 / hbm, 'note': note}
        if cpu:
            e['cpu_oracle'] = cpu
            e['speedup_vs_cpu_oracle'] = e['units_per_s'] / cpu['units_per_s']
        res['kernels'][key] = e

    S = ow.SortedAdj.from_csr(g.rowptr_host, g.col_host)

    def cpu_rate(fn, count_of):
        """runs fn(i) for i = 0, 1, ... for ~cpu_seconds; returns units/s of the pure-Python oracle (1 core)."""
        t0, n, units = time.perf_counter(), 0, 0
        while time.perf_counter() - t0 < cpu_seconds:
            units += count_of(fn(n))
            n += 1
        dt = time.perf_counter() - t0
        return {'units_per_s': units / dt, 'cores': 1, 'sample': '%d calls, %.1f s' % (n, dt)}

    # ---- a1/a2: structure anchor patches = triangular random walks over the whole graph ----
    n_bench = max(P_tot, 20000)                      # the data set needs P_tot walks (a few hundred): bench a launch that fills the GPU
    t, patches_big = timed(lambda: ops.walk_full(g, n_bench, Ls, beta, 1), reps, flush)
    steps = int((patches_big != 0).sum().item())
    visited = patches_big[patches_big != 0].long() - 1
    mean_deg = float(deg[visited].mean().item())
    cpu = cpu_rate(lambda i: ow.sample_structure_anchor_patches(S, 2, Ls, beta, ow.philox_patch_factory(100 + i)), lambda p: int((np.asarray(p) != 0).sum()))
    rec('walk_full', t, steps, 'walk steps', steps * (8 * mean_deg + 16), 'warp per walk; %d walks of length <= %d; per step 4(deg(prev)+deg(curr))+16 B, '
        'mean degree of visited nodes %.0f' % (n_bench, Ls, mean_deg), cpu)
    patches = ops.walk_full(g, P_tot, Ls, beta, 1)
    for border in (False, True):
        reps_p = max(1, 4000 // P_tot)
        big = patches.repeat(reps_p, 1)
        t, w = timed(lambda: ops.walk_patch(g, big, W, T, beta, border, 2), reps, flush)
        steps = int((w != 0).sum().item())
        cpu = cpu_rate(lambda i: ow.perform_random_walks(S, patches[i % P_tot:i % P_tot + 1].cpu().numpy().astype(np.int64), W, T, beta, not border,
                                                         ow.philox_walk_factory(7, W)), lambda p: int((np.asarray(p) != 0).sum()))
        rec('walk_patch_border' if border else 'walk_patch_inside', t, steps, 'walk steps', steps * (8 * mean_deg + 16),
            'CTA per patch (%d patches x %d walks x %d steps), neighbour lists filtered against the patch' % (big.shape[0], W, T), cpu)
    # ---- components, border sets, anchor draws ----
    cc = prepare.initialize_cc_ids(g, subs['train'])
    n_sub, C, Lcc = cc.shape
    rp, ri = ragged_from_padded(cc.reshape(n_sub * C, Lcc))
    rptr, ritems = torch.from_numpy(rp).to(dev), torch.from_numpy(ri).to(dev)
    k = hp['neigh_sample_border_size']
    t, (bptr, bitems) = timed(lambda: ops.border_khop(g, rptr, ritems, k), reps, flush)
    n_cc = int((cc[:, :, 0] != 0).sum())
    sum_deg = float(deg[ritems.long() - 1].sum().item())
    g_mean_deg = float(g.col.numel()) / g.n_nodes
    rec('border_khop', t, n_cc, 'components', 4.0 * bitems.numel() + 4.0 * sum_deg * (1 + (g_mean_deg if k == 2 else 0)),
        '%d-hop border sets, %d ids out; neighbour lists of the members (and, for k = 2, of their neighbours: estimated with the mean degree)' % (k, bitems.numel()))
    width_b = int((bptr[1:] - bptr[:-1]).max().item())
    A = hp['n_anchor_patches_N_out']
    t, _ = timed(lambda: ops.sample_rows(bptr.to(torch.int32), bitems, width_b, A, True, 0, 1, False), reps, flush)
    rec('sample_rows_N_border', t, n_sub * C * A, 'draws', n_sub * C * A * 12.0, 'thread per draw')
    # ---- gamma: hop table, min-hop similarity, degree sequences, DTW ----
    t, hop = timed(lambda: ops.hop_table(g), 1, flush)
    rec('hop_table', t, g.n_nodes, 'BFS sources', g.n_nodes * (2.0 * g.col.numel() * 4 / 2 + g.n_nodes), 'CTA per source, bitmap frontier; N^2 uint8 out')
    ids = ops.sample_rows(bptr.to(torch.int32), bitems, width_b, A, True, 0, 1, False)
    t, _ = timed(lambda: ops.sp_min_gather(hop, rptr.to(torch.int32), ritems, ids.contiguous()), reps, flush)
    mean_len = float(ritems.numel()) / max(n_cc, 1)
    rec('sp_min_gather', t, n_sub * C * A, '(component, anchor) pairs', n_cc * A * mean_len * 32.0,
        '1-byte gathers from the hop table: one 32 B sector per (member, anchor); useful bytes = 1/32 of that (sector efficiency 3 %)')
    n_rows = min(n_sub * C, 2048)
    t, _ = timed(lambda: ops.sp_min_dense(hop, rptr[:n_rows + 1].to(torch.int32), ritems), reps, flush)
    rows_len = float(rp[n_rows])
    rec('sp_min_dense', t, n_rows, 'component rows', rows_len * g.n_nodes + n_rows * 4.0 * g.n_nodes, 'streams |cc| hop-table rows, writes fp32 (N)')
    flat = torch.from_numpy(cc.reshape(n_sub * C, Lcc)).to(dev)
    for internal in (True, False):
        t, (sa, la) = timed(lambda: ops.degree_seq(g, flat, internal), reps, flush)
        rec('degree_seq_' + ('internal' if internal else 'border'), t, n_sub * C, 'rows', 16.0 * ritems.numel() + (4.0 * sum_deg if internal else 0.0),
            'warp per row: ids + CSR offsets in, sorted degrees out' + (' + the members\' neighbour lists for the induced-degree count' if internal else ''))
    sa, la = ops.degree_seq(g, flat, True)
    sb, lb = ops.degree_seq(g, patches, True)
    t, sims = timed(lambda: ops.dtw_batch(sa, la, sb, lb, ops.DTW_FASTDTW_R1, max_len_a=Lcc, max_len_b=patches.shape[1]), reps, flush)
    pairs = n_cc * P_tot
    sa_h, la_h, sb_h, lb_h = sa.cpu().numpy(), la.cpu().numpy(), sb.cpu().numpy(), lb.cpu().numpy()
    rows_valid = np.nonzero(la_h)[0]
    cpu = cpu_rate(lambda i: og.calc_dtw(sa_h[rows_valid[i % len(rows_valid)], :la_h[rows_valid[i % len(rows_valid)]]].tolist(),
                                         sb_h[i % P_tot, :lb_h[i % P_tot]].tolist()), lambda v: 1)
    rec('dtw_batch_fastdtw_r1', t, pairs, '(component, patch) pairs', pairs * 4.0 * (float(la_h[rows_valid].mean()) + float(lb_h.mean()) + 1),
        'thread per pair, fp64 DP; FP64-issue / shared-memory bound, not HBM: mean lengths %.1f x %.1f' % (float(la_h[rows_valid].mean()), float(lb_h.mean())), cpu)
    # exact DTW on both mappings: DP cells/s (n*m per pair) — the sub-warp wavefront (registers + shuffles) against one thread per pair
    cells = float(la_h[rows_valid].astype(np.float64).sum()) * float(lb_h.astype(np.float64).sum())
    for mode, key, note, bk in ((ops.DTW_EXACT, 'dtw_batch_exact_wavefront', 'components bucketed by length (<= 2: thread per pair; else G = 4 ... 32 '
                                 'lanes per pair), DP columns in registers, __shfl_up_sync hand-off, patch sequence staged in smem', True),
                                (ops.DTW_EXACT, 'dtw_batch_exact_wavefront_one_launch', 'one launch, G from the longest component', False),
                                (ops.DTW_EXACT_THREAD, 'dtw_batch_exact_thread', 'thread per pair, rolling fp64 rows in smem', False)):
        t, sims_e = timed(lambda: ops.dtw_batch(sa, la, sb, lb, mode, max_len_a=Lcc, max_len_b=patches.shape[1], bucketed=bk), reps, flush)
        rec(key, t, pairs, '(component, patch) pairs', pairs * 4.0 * (float(la_h[rows_valid].mean()) + float(lb_h.mean()) + 1),
            note + '; %.3g DP cells/s; longest component %d' % (cells / t, int(la_h.max())))
    return res


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--workloads', nargs='+', default=['ppi_bp', 'hpo_metab', 'em_user'])
    ap.add_argument('--cpu-seconds', type=float, default=4.0)
    a = ap.parse_args()
    out = [run(w, a.cpu_seconds) for w in a.workloads]
    print(json.dumps(out, indent=1))
