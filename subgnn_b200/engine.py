"""Step engine: device-resident tables + one flat parameter arena + the fused CUDA step.

Host-side orchestration of the kernels behind include/subgnn_b200.h for the per-step path
(SubGNN.py:225-348 forward / training_step, :1156-1164 Adam + backward, Lightning's gradient clipping).
torch provides device memory, streams, CUDA-graph capture and (for data parallel) the NCCL allreduce;
all arithmetic happens in libsubgnn_b200.so.

Data layout in HBM:
  * parameters / gradients / Adam m, v: four flat fp32 arenas with identical offsets; nn.Parameters of the
    SubGNN module are views into the parameter arena (reference state_dict names and shapes);
  * per split: ragged component table (valid components only) + per-layer anchor ids and RESOLVED
    similarities (component x sampled anchor) instead of the reference's dense (n_sub, C, N) slab;
  * per step: only the batch's subgraph indices travel host -> device.
"""
import ctypes as C
import os
import math
import sys

import numpy as np
import torch

from . import _abi
from ._abi import ModelDesc, call, ptr

CHANNELS = ('neighborhood', 'position', 'structure')
SIDES = ('internal', 'border')


def _flag(env, default):
    """tuning switch: environment variable (0/1) over the hyper-parameter / built-in default."""
    v = os.environ.get(env)
    return bool(default) if v is None else v not in ('0', '', 'false')


def _align(n, a=4):
    return (n + a - 1) // a * a


# ------------------------------------------------------------------------------------------------------
class ParamArena:
    """Flat fp32 storage for every trainable tensor, laid out for the kernels (see header comment of
    subgnn_model_desc) and exposed under the reference's state_dict names."""

    def __init__(self, hp, n_nodes, num_classes, hid_dim, n_train=0, C_pad=0, device='cuda', alloc=None):
        D, L = hp['node_embed_size'], hp['n_layers']
        self.D, self.L = D, L
        self.device = torch.device(device)
        self.entries = {}      # name -> (offset, shape)
        off = 0

        def add(name, shape):
            nonlocal off
            n = int(np.prod(shape))
            self.entries[name] = (off, tuple(shape))
            off += n

        def pad():
            nonlocal off
            off = _align(off)

        self.embed_trainable = not hp['freeze_node_embeds']
        if self.embed_trainable:
            add('node_embeddings.weight', (n_nodes + 1, D)); pad()
        self.mpn_base = {}
        blk = 2 * D * D + 2 * D + 1
        self.mpn_blk = blk
        for ch, key in zip(CHANNELS, ('use_neighborhood', 'use_position', 'use_structure')):
            if not hp[key]:
                continue
            pad()
            self.mpn_base[ch] = off
            for l in range(L):
                for side in SIDES:
                    pre = '%s_mpns.%d.%s.' % (ch, l, side)
                    add(pre + 'linear.weight', (D, 2 * D))
                    add(pre + 'linear.bias', (D,))
                    add(pre + 'linear_position.weight', (1, D))
                    add(pre + 'linear_position.bias', (1,))
            pad()
        H = D
        self.lstm_layers = hp['lstm_n_layers']
        self.lstm_off = {}
        for k in range(self.lstm_layers):
            din = D if k == 0 else 2 * H
            pad()
            self.lstm_off[k] = {}
            for nm, shape in (('weight_ih', (4 * H, din)), ('weight_hh', (4 * H, H)), ('bias_ih', (4 * H,)), ('bias_hh', (4 * H,))):
                pad()
                self.lstm_off[k][nm] = off
                add('lstm.lstm.%s_l%d' % (nm, k), shape)
                add('lstm.lstm.%s_l%d_reverse' % (nm, k), shape)      # both directions contiguous
        pad()
        add('lstm.linear.weight', (D, 2 * H)); pad()
        add('lstm.linear.bias', (D,)); pad()
        for nm, shape in (('lin', (hp['linear_hidden_dim_1'], hid_dim)), ('lin2', (hp['linear_hidden_dim_2'], hp['linear_hidden_dim_1'])),
                          ('lin3', (num_classes, hp['linear_hidden_dim_2']))):
            add(nm + '.weight', shape); pad()
            add(nm + '.bias', (shape[0],)); pad()
        self.cc_tables = bool(hp['trainable_cc'])
        if self.cc_tables:
            for nm in ('N_I', 'N_B', 'S_I', 'S_B', 'P_I', 'P_B'):
                add('train_%s_cc_embed' % nm, (n_train, C_pad, D)); pad()
        self.size = _align(off)
        z = lambda: torch.zeros(self.size, dtype=torch.float32, device=self.device)
        # alloc: allocator of the two arenas that peers map (data-parallel exchange over NVLink peer memory, DpExchange)
        zs = (lambda: alloc(self.size).zero_()) if alloc is not None else z
        self.params, self.grads, self.m, self.v = zs(), zs(), z(), z()

    def view(self, name, which='params'):
        off, shape = self.entries[name]
        return getattr(self, which)[off:off + int(np.prod(shape))].view(shape)

    def addr(self, name, which='params'):
        off, _ = self.entries[name]
        return getattr(self, which).data_ptr() + 4 * off

    def base_addr(self, off, which='params'):
        return getattr(self, which).data_ptr() + 4 * off

    def load_state_dict(self, sd):
        for name in self.entries:
            if name in sd:
                self.view(name).copy_(torch.as_tensor(sd[name]).to(self.device, torch.float32).reshape(self.entries[name][1]))

    def state_dict(self):
        return {k: self.view(k).detach().clone() for k in self.entries}


# ------------------------------------------------------------------------------------------------------
class SplitTables:
    """Ragged, device-resident form of one split (train / val / test) of the prepared data."""

    def __init__(self, prepared, split, hp, device, graph=None):
        dev = torch.device(device)
        L = hp['n_layers']
        cc = np.asarray(prepared['cc_ids'][split])
        n_sub, C_pad, Lcc = cc.shape
        valid = cc[:, :, 0] != 0
        counts = valid.sum(axis=1)
        assert (valid[:, :1].all() or n_sub == 0), 'every subgraph needs at least one component'
        # valid components come first along C (SubGNN.py:596-599 appends the padding)
        assert all(valid[s, :counts[s]].all() for s in range(n_sub))
        self.n_sub, self.C_pad = n_sub, C_pad
        self.n_cc = int(counts.sum())
        sub_ccptr = np.zeros(n_sub + 1, dtype=np.int64)
        np.cumsum(counts, out=sub_ccptr[1:])
        rows = cc[valid]                                   # (n_cc, Lcc) in (sub, c) order
        lens = (rows != 0).sum(axis=1)
        nodeptr = np.zeros(self.n_cc + 1, dtype=np.int64)
        np.cumsum(lens, out=nodeptr[1:])
        nodes = rows[rows != 0]
        row_sub = np.repeat(np.arange(n_sub), counts)
        sub_maxlen = np.zeros(n_sub, dtype=np.int64)
        np.maximum.at(sub_maxlen, row_sub, lens)
        self.max_cc_per_sub = int(counts.max()) if n_sub else 1
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        tf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        self.sub_ccptr, self.sub_maxlen = t32(sub_ccptr), t32(sub_maxlen)
        self.cc_nodeptr, self.cc_nodes = t32(nodeptr), t32(nodes)
        self.row_sub = t32(row_sub)
        self.counts_host = counts
        lab = np.asarray(prepared['labels'][split])
        self.multilabel = bool(prepared.get('multilabel', False))
        self.labels = tf(lab) if self.multilabel else t32(lab.reshape(-1))
        dense = prepared.get('NP_sim')
        dense_rows = np.asarray(dense[split])[valid] if (dense is not None and split in dense) else None     # (n_cc, N)
        hop = graph.hop if (graph is not None and dense_rows is None) else None

        def resolve(ids_rows, anchor_row=None):
            """similarity of every (component row, anchor) pair; ids_rows: (n_lists, A) int ids."""
            if dense_rows is not None:
                lists = ids_rows if anchor_row is None else ids_rows[anchor_row]
                g = np.take_along_axis(dense_rows, np.maximum(lists - 1, 0).astype(np.int64), axis=1)
                return np.where(lists > 0, g, 0).astype(np.float32)
            from . import ops
            assert hop is not None, 'need either prepared["NP_sim"] or a hop table on the graph'
            ar = None if anchor_row is None else t32(anchor_row)
            return ops.sp_min_gather(hop, self.cc_nodeptr, self.cc_nodes, t32(ids_rows), ar)

        as_dev = lambda x: x if isinstance(x, torch.Tensor) else tf(x)
        self.n_ids, self.n_sim = [None, None], [None, None]
        if hp['use_neighborhood']:
            for si, key in enumerate(('anchors_neigh_int', 'anchors_neigh_border')):
                ids = [np.asarray(prepared[key][split][l])[valid] for l in range(L)]
                self.n_ids[si] = t32(np.stack(ids))
                self.n_sim[si] = torch.stack([as_dev(resolve(i)) for i in ids]).contiguous()
        self.p_int_ids = self.p_bor_ids = None
        self.p_sim = [None, None]
        if hp['use_position']:
            pi = [np.asarray(prepared['anchors_pos_int'][split][l]) for l in range(L)]
            pb = [np.asarray(prepared['anchors_pos_ext'][l]).reshape(1, -1) for l in range(L)]
            self.p_int_ids, self.p_bor_ids = t32(np.stack(pi)), t32(np.concatenate(pb))
            self.p_sim[0] = torch.stack([as_dev(resolve(i, row_sub)) for i in pi]).contiguous()
            self.p_sim[1] = torch.stack([as_dev(resolve(i, np.zeros(self.n_cc, dtype=np.int64))) for i in pb]).contiguous()
        self.s_sim = [None, None]
        if hp['use_structure']:
            for si, key in enumerate(('I_S_sim', 'B_S_sim')):
                tab = np.asarray(prepared[key][split])[valid]                      # (n_cc, P_tot)
                self.s_sim[si] = tf(np.stack([tab[:, np.asarray(prepared['anchors_structure'][l][1], dtype=np.int64)] for l in range(L)]))


# ------------------------------------------------------------------------------------------------------
class LstmRunner:
    """Sequences the LSTM kernels for the structure anchor patches (all layers / sides in one batch:
    the reference shares one LSTM across them, SubGNN.py:175)."""

    def __init__(self, arena, hp, walks, n_groups, device, n_seq=None, T=None):
        """walks: int32 device tensor (n_groups * W, T), or None for dense inputs of shape (n_seq, T, D)
        (module-level LSTM.forward, SubGNN.py:76-88)."""
        self.arena, self.hp = arena, hp
        self.dev = torch.device(device)
        self.D = self.H = hp['node_embed_size']
        self.nl = hp['lstm_n_layers']
        self.W = hp['n_triangular_walks']
        self.n_groups = n_groups
        self.walks = walks.contiguous() if walks is not None else None
        self.n_seq, self.T = walks.shape if walks is not None else (n_seq, T)
        self.sum_mode = 1 if hp['lstm_aggregator'] == 'sum' else 0
        if hp['lstm_aggregator'] not in ('sum', 'last'):
            raise NotImplementedError(hp['lstm_aggregator'])                       # SubGNN.py:86-87
        self.p_drop = float(hp['lstm_dropout']) if self.nl > 1 else 0.0
        H, M = self.H, self.n_seq * self.T
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.dev)
        self.G = [z(M, 8 * H) for _ in range(self.nl)]
        self.OUT = [z(M + 1, 2 * H) for _ in range(self.nl)]
        self.CS = [z(M, 2 * H) for _ in range(self.nl)]
        self.dOUT = [z(M, 2 * H) for _ in range(self.nl)]
        self.X = [None] + [z(M + 1, 2 * H) for _ in range(1, self.nl)] if self.p_drop > 0 else [None] * self.nl
        self.whh_t = [z(2 * H * 4 * H) for _ in range(self.nl)]
        self.bsum = [z(8 * H) for _ in range(self.nl)]
        # head: the walks of a patch are summed BEFORE the (linear) head, so it runs on one row per patch (n_groups rows)
        self.AGG = z(n_groups, 2 * H)
        self.EMB, self.dEMB = z(n_groups, self.D), z(n_groups, self.D)
        self._aux = None
        t = torch.arange(M, device=self.dev, dtype=torch.int64)
        tt = t % self.T
        zero_row = torch.full_like(t, M)
        self.hprev = [torch.where(tt > 0, t - 1, zero_row).to(torch.int32).contiguous(),
                      torch.where(tt < self.T - 1, t + 1, zero_row).to(torch.int32).contiguous()]
        self.ids_flat = self.walks.reshape(-1).contiguous() if walks is not None else None
        self.ids_last = self.walks[:, -1].contiguous() if walks is not None else None
        self.dense_x = None
        # tensor-core (tcgen05, 3xTF32) projections need 16-byte aligned rows; otherwise the FFMA tiles are used
        self.use_tc = (self.D % 4 == 0) and hp.get('b200_tensor_core_gemm', True)
        # TMA-fed warp-specialised grouped GEMM (tcgemm_ws.cu): dense operands only, so the walk-node rows of the embedding table are
        # gathered ONCE per step into X0 (they feed the layer-0 projection and its weight gradient); the gate-gradient consumers
        # of a layer (input gradient + weight gradients) go out as grouped launches.  SUBGNN_TC_LEGACY=1 / SUBGNN_GEMM_GROUPED=0
        # keep the round-1 per-GEMM launches for A/B runs.
        self.ws = self.use_tc and bool(_abi.lib.subgnn_tc_ws_available()) and not _flag('SUBGNN_TC_LEGACY', False) and hp.get('b200_ws_gemm', True)
        self.grouped = int(os.environ.get('SUBGNN_GEMM_GROUPED', hp.get('b200_gemm_grouped', 1))) if self.ws else 0
        # top layer of a multi-layer stack: the all-rows and the last-step-rows input-gradient products in ONE launch (atomics onto a
        # zero-filled dOUT) instead of two dependent launches on the chain.  Measured on B200 (PPI-BP shape, same box): 0.3357 ms/step
        # with the merge against 0.3224 without — the vector atomics of the 79-tile product cost more than the 7 us launch they
        # save — so it is a switch, off.  Likewise SUBGNN_LSTM_PREP_BESIDE (weight packing on a branch beside the gather): 0.3357
        # against 0.3148.
        self.bi_merge = _flag('SUBGNN_BI_MERGE', hp.get('b200_bi_merge', False))
        self.prep_beside = _flag('SUBGNN_LSTM_PREP_BESIDE', hp.get('b200_lstm_prep_beside', False))
        # top layer with the 'last' aggregator: the input gradient of its reverse direction (last-step rows only) is STORED into a side
        # buffer by the same grouped launch as the all-rows product and added by the next BPTT on load (subgnn_lstm_recur_bwd_add),
        # instead of a second, dependent launch that read-add-stores onto the first one's rows
        self.bi_side = _flag('SUBGNN_BI_SIDE', hp.get('b200_bi_side', True)) and bool(_abi.lib.subgnn_lstm_fused_dropout_supported(self.H))
        self.dOUT_side = z(M, 2 * H) if (self.bi_side and self.nl > 1) else None      # only its rows t = T-1 are ever written / read
        self.X0 = z(M, self.D) if (self.ws and walks is not None) else None
        self.fwd_fn = 'subgnn_tc_linear_fwd' if self.use_tc else 'subgnn_linear_fwd'
        self.bwi_fn = 'subgnn_tc_linear_bwd_input' if self.use_tc else 'subgnn_linear_bwd_input'
        # inter-layer dropout fused into the recurrence kernels (mask written / applied in place of two element-wise launches per layer)
        self.fused_drop = bool(_abi.lib.subgnn_lstm_fused_dropout_supported(self.H)) and hp.get('b200_fused_lstm_dropout', True)
        # 'last' aggregator: the head's gradient reaches the top layer's output only at t = T-1; the head then writes those rows
        # alone and the BPTT kernel takes every other row of dOUT as zero without reading it (include/subgnn_b200.h flags)
        self.last_only = (not self.sum_mode) and bool(_abi.lib.subgnn_lstm_fused_dropout_supported(self.H)) and _flag('SUBGNN_LSTM_LAST_ONLY', True)

    def steps(self, k):
        top = k == self.nl - 1
        return (self.T, 1) if (top and not self.sum_mode) else (self.T, self.T)

    # Independent launches (the reverse-direction projection of the top layer's last rows, every weight-gradient GEMM of the
    # backward pass) go to two auxiliary streams, forked from / joined into the stream the runner is called on with events;
    # under graph capture they become parallel branches of the step graph.
    def _aux_streams(self):
        if self._aux is None:
            self._aux = [torch.cuda.Stream(device=self.dev) for _ in range(3)]
        return self._aux

    def forward(self, E_ptr, training, seed, step_dev, st, dense_x=None, step_event=None):
        """step_event: recorded by the caller once *step_dev holds this step's counter (fused step: the chain is forked BEFORE the
        counter kernel; only the recurrences with fused dropout / the dropout kernels read it)."""
        a, H, D, M, T = self.arena, self.H, self.D, self.n_seq * self.T, self.T
        cur = torch.cuda.current_stream()
        aux = self._aux_streams()
        self.dense_x = dense_x
        if dense_x is not None:
            E_ptr = ptr(dense_x)
        gather = self.X0 is not None and dense_x is None
        # weight packing (W_hh^T, b_ih + b_hh) of every layer beside the once-per-step embedding gather, not in front of it
        ps = aux[0] if (gather and self.prep_beside) else cur
        if ps is not cur:
            ps.wait_stream(cur)
        if self.nl > 1 and self.nl <= 8:            # every layer's packing in one launch
            if getattr(self, '_prep_tabs', None) is None:
                tab = lambda f: (C.c_ulonglong * self.nl)(*[int(f(k)) for k in range(self.nl)])
                self._prep_tabs = (tab(lambda k: a.base_addr(a.lstm_off[k]['weight_hh'])), tab(lambda k: a.base_addr(a.lstm_off[k]['bias_ih'])),
                                   tab(lambda k: a.base_addr(a.lstm_off[k]['bias_hh'])), tab(lambda k: self.whh_t[k].data_ptr()),
                                   tab(lambda k: self.bsum[k].data_ptr()))
            call('subgnn_lstm_prep_layers', *self._prep_tabs, self.nl, H, ps.cuda_stream)
        else:
            for k in range(self.nl):
                o = a.lstm_off[k]
                call('subgnn_lstm_prep', a.base_addr(o['weight_hh']), a.base_addr(o['bias_ih']), a.base_addr(o['bias_hh']),
                     ptr(self.whh_t[k]), ptr(self.bsum[k]), H, ps.cuda_stream)
        if gather:
            call('subgnn_gather_rows', E_ptr, ptr(self.ids_flat), ptr(self.X0), M, D, st)     # anchor_patch_samplers.py:409, once per step
        if ps is not cur:
            cur.wait_stream(ps)
        for k in range(self.nl):
            o = a.lstm_off[k]
            fused = self.fused_drop and self.p_drop > 0 and training
            x_ptr, ldx, ids, din = self._layer_input(k, E_ptr, training, seed, step_dev, st, make=not fused)
            sf, sr = self.steps(k)
            w_ih, G = a.base_addr(o['weight_ih']), ptr(self.G[k])
            if self.ws and ids is None:
                descs = [_abi.gemm_desc(_abi.GEMM_FWD, x_ptr, ldx, w_ih, din, G, 8 * H, M, 8 * H if sr == T else 4 * H, din, bias=ptr(self.bsum[k]))]
                if sr != T:      # 'last' aggregator, top layer: reverse-direction gates of the rows t = T-1 only (SubGNN.py:83), same launch
                    xl, ldxl, _ = self._last_rows(k, x_ptr, ldx, None)
                    descs.append(_abi.gemm_desc(_abi.GEMM_FWD, xl, ldxl, w_ih + 4 * (4 * H * din), din, G + 4 * (((T - 1) * 2 + 1) * 4 * H), T * 8 * H,
                                                self.n_seq, 4 * H, din, bias=self.bsum[k].data_ptr() + 4 * 4 * H))
                _abi.gemm_group(descs, st)
            elif sr == T:
                call(self.fwd_fn, x_ptr, ldx, ids, w_ih, din, ptr(self.bsum[k]), G, 8 * H, M, 8 * H, din, 0, st)
            else:
                # 'last' aggregator, top layer: the reverse direction is only ever read at t = T-1 (SubGNN.py:83), so its
                # input projection is computed for those n_seq rows only (disjoint gate columns: runs beside the main GEMM)
                aux[0].wait_stream(cur)
                with torch.cuda.stream(aux[0]):
                    xl, ldxl, idsl = self._last_rows(k, x_ptr, ldx, ids)
                    call(self.fwd_fn, xl, ldxl, idsl, w_ih + 4 * (4 * H * din), din, self.bsum[k].data_ptr() + 4 * 4 * H,
                         G + 4 * (((T - 1) * 2 + 1) * 4 * H), T * 8 * H, self.n_seq, 4 * H, din, 0, aux[0].cuda_stream)
                call(self.fwd_fn, x_ptr, ldx, ids, w_ih, din, ptr(self.bsum[k]), G, 8 * H, M, 4 * H, din, 0, st)
                cur.wait_stream(aux[0])
            if k == 0 and step_event is not None:
                cur.wait_event(step_event)           # the gather and the first projection do not read the step counter; everything below may
            if fused and k + 1 < self.nl:            # also writes X[k+1] = dropout(OUT[k]), the next layer's input
                call('subgnn_lstm_recur_fwd_drop', G, ptr(self.whh_t[k]), ptr(self.OUT[k]), ptr(self.CS[k]), self.n_seq, T, H, sf, sr,
                     ptr(self.X[k + 1]), self.p_drop, seed, 8 + k + 1, step_dev, st)
            else:
                call('subgnn_lstm_recur_fwd', G, ptr(self.whh_t[k]), ptr(self.OUT[k]), ptr(self.CS[k]), self.n_seq, T, H, sf, sr, st)
        call('subgnn_lstm_head_fwd', ptr(self.OUT[-1]), ptr(self.AGG), ptr(self.EMB), a.addr('lstm.linear.weight'), a.addr('lstm.linear.bias'),
             self.n_groups, self.W, T, 2 * H, D, self.sum_mode, st)

    def _wgrad(self, dy, ldy, x, ldx, ids, dw, lddw, db, M, N, K, st):
        if self.use_tc and M >= 256:
            call('subgnn_tc_linear_bwd_weight', dy, ldy, x, ldx, ids, dw, lddw, db, M, N, K, st)
        else:
            call('subgnn_linear_bwd_weight', dy, ldy, x, ldx, ids, dw, lddw, db, M, N, K, None, st)

    def _layer_input(self, k, E_ptr, training, seed, step_dev, st, make=True):
        """(x pointer, leading dim, gather ids, K) of layer k's input rows (all n_seq*T of them)."""
        H, D, M = self.H, self.D, self.n_seq * self.T
        if k == 0:
            if self.X0 is not None and self.dense_x is None:
                return ptr(self.X0), D, None, D
            return E_ptr, D, (ptr(self.ids_flat) if self.dense_x is None else None), D
        x = self.OUT[k - 1]
        if self.p_drop > 0 and training:
            if make:
                call('subgnn_dropout', ptr(x), ptr(self.X[k]), M * 2 * H, self.p_drop, seed, 8 + k, step_dev, st)
            x = self.X[k]
        return ptr(x), 2 * H, None, 2 * H

    def _last_rows(self, k, x_ptr, ldx, ids):
        """the n_seq input rows at t = T-1: strided view of a dense input, or the last id of every walk."""
        T = self.T
        if ids is not None:
            return x_ptr, ldx, ptr(self.ids_last)
        return x_ptr + 4 * ((T - 1) * ldx), T * ldx, None

    def backward(self, E_ptr, dE_ptr, training, seed, step_dev, st, dense_dx=None):
        """consumes self.dEMB; accumulates into the gradient arena (and dE, or writes dense_dx for dense inputs).
        Critical chain on the calling stream: head input gradient -> [recurrence BPTT -> input gradient -> dropout mask] per
        layer; every weight-gradient GEMM only consumes what the chain has already produced and runs on the auxiliary streams."""
        a, H, D, M, T = self.arena, self.H, self.D, self.n_seq * self.T, self.T
        cur = torch.cuda.current_stream()
        aux = self._aux_streams()
        dense = self.dense_x is not None
        if dense:
            E_ptr = ptr(self.dense_x)
        g = 'grads'
        aux[1].wait_stream(cur)
        with torch.cuda.stream(aux[1]):
            if self.ws and self.bi_merge and self.nl > 1 and self.steps(self.nl - 1)[1] != T:
                # the top layer's two input-gradient products (all rows x forward gates, last-step rows x reverse gates) add into
                # dOUT[nl-2] with atomics in one grouped launch: its zero fill runs here, beside the head gradient
                call('subgnn_fill_zero', ptr(self.dOUT[self.nl - 2]), self.dOUT[self.nl - 2].numel(), aux[1].cuda_stream)
            call('subgnn_linear_bwd_weight', ptr(self.dEMB), D, ptr(self.AGG), 2 * H, None, a.addr('lstm.linear.weight', g), 2 * H,
                 None, self.n_groups, D, 2 * H, None, aux[1].cuda_stream)
        call('subgnn_lstm_head_bwd', ptr(self.dEMB), a.addr('lstm.linear.weight'), ptr(self.dOUT[-1]), a.addr('lstm.linear.bias', g),
             self.n_groups, self.W, T, 2 * H, D, 2 if self.last_only else self.sum_mode, st)
        side_for = -1                                # layer whose BPTT adds the side buffer
        for k in range(self.nl - 1, -1, -1):
            o = a.lstm_off[k]
            sf, sr = self.steps(k)
            full = sr == T
            # the bias gradients (d b_ih == d b_hh == column sums of dG) are accumulated inside the recurrence kernel
            fused = self.fused_drop and self.p_drop > 0 and training
            flags = (1 if full else 0) | (2 if (self.last_only and k == self.nl - 1) else 0)
            if side_for == k:                        # two gradient sources: dOUT[k] + the side buffer's rows t = T-1 (written by layer k+1)
                masked = fused and k + 1 < self.nl
                call('subgnn_lstm_recur_bwd_add', ptr(self.G[k]), a.base_addr(o['weight_hh']), ptr(self.OUT[k]), ptr(self.CS[k]),
                     ptr(self.dOUT[k]), ptr(self.dOUT_side), self.n_seq, T, H, sf, sr, flags, a.base_addr(o['bias_ih'], g),
                     a.base_addr(o['bias_hh'], g), self.p_drop if masked else 0.0, seed, 8 + k + 1, step_dev if masked else None, st)
            elif fused and k + 1 < self.nl:          # dOUT[k] is the gradient w.r.t. X[k+1] = dropout(OUT[k]): mask applied on load
                call('subgnn_lstm_recur_bwd_drop', ptr(self.G[k]), a.base_addr(o['weight_hh']), ptr(self.OUT[k]), ptr(self.CS[k]),
                     ptr(self.dOUT[k]), self.n_seq, T, H, sf, sr, flags, a.base_addr(o['bias_ih'], g),
                     a.base_addr(o['bias_hh'], g), self.p_drop, seed, 8 + k + 1, step_dev, st)
            else:
                call('subgnn_lstm_recur_bwd', ptr(self.G[k]), a.base_addr(o['weight_hh']), ptr(self.OUT[k]), ptr(self.CS[k]), ptr(self.dOUT[k]),
                     self.n_seq, T, H, sf, sr, flags, a.base_addr(o['bias_ih'], g), a.base_addr(o['bias_hh'], g), st)
            dG = ptr(self.G[k])
            x_ptr, ldx, ids, din = self._layer_input(k, E_ptr, training, seed, step_dev, st, make=False)
            w_ih, gw_ih = a.base_addr(o['weight_ih']), a.base_addr(o['weight_ih'], g)
            n_out = 8 * H if full else 4 * H                      # gate columns that carry gradient on every row
            dG_last = dG + 4 * (((T - 1) * 2 + 1) * 4 * H)         # reverse-direction gates of the rows t = T-1
            if self.ws and ids is None:
                side = None
                if self.dOUT_side is not None and k > 0 and not full and not (self.p_drop > 0 and training and not fused):
                    side, side_for = ptr(self.dOUT_side), k - 1
                self._backward_gemms_ws(k, dG, dG_last, x_ptr, ldx, din, full, n_out, dE_ptr, dense, dense_dx, st, cur, aux, side)
                if k > 0 and self.p_drop > 0 and training and not fused:
                    call('subgnn_dropout', ptr(self.dOUT[k - 1]), ptr(self.dOUT[k - 1]), M * 2 * H, self.p_drop, seed, 8 + k, step_dev, st)
                continue
            # ---- weight gradients: off the critical chain ----
            for a_ in aux:
                a_.wait_stream(cur)
            with torch.cuda.stream(aux[0]):
                s0 = aux[0].cuda_stream
                self._wgrad(dG, 8 * H, x_ptr, ldx, ids, gw_ih, din, None, M, n_out, din, s0)
                if not full:
                    xl, ldxl, idsl = self._last_rows(k, x_ptr, ldx, ids)
                    self._wgrad(dG_last, T * 8 * H, xl, ldxl, idsl, gw_ih + 4 * (4 * H * din), din, None, self.n_seq, 4 * H, din, s0)
            for d_ in range(2 if full else 1):                     # reverse direction took one step from h = 0: no W_hh gradient
                with torch.cuda.stream(aux[1 + d_]):
                    self._wgrad(dG + 4 * (d_ * 4 * H), 8 * H, self.OUT[k].data_ptr() + 4 * (d_ * H), 2 * H, ptr(self.hprev[d_]),
                                a.base_addr(o['weight_hh'], g) + 4 * (d_ * 4 * H * H), H, None, M, 4 * H, H, aux[1 + d_].cuda_stream)
            # ---- input gradient: the chain ----
            if k > 0:
                dx_ptr, lddx, scat = ptr(self.dOUT[k - 1]), 2 * H, None
            elif dense:
                if dense_dx is None:
                    continue
                dx_ptr, lddx, scat = ptr(dense_dx), D, None
            elif dE_ptr:
                dx_ptr, lddx, scat = dE_ptr, D, ptr(self.ids_flat)
            else:
                continue
            acc = 1 if scat is not None else 0
            call(self.bwi_fn, dG, 8 * H, w_ih, din, dx_ptr, lddx, scat, M, n_out, din, acc, st)
            if not full:
                if scat is not None:
                    call(self.bwi_fn, dG_last, T * 8 * H, w_ih + 4 * (4 * H * din), din, dx_ptr, lddx, ptr(self.ids_last),
                         self.n_seq, 4 * H, din, 1, st)
                else:
                    call(self.bwi_fn, dG_last, T * 8 * H, w_ih + 4 * (4 * H * din), din, dx_ptr + 4 * ((T - 1) * lddx), T * lddx,
                         None, self.n_seq, 4 * H, din, 1, st)
            if k > 0 and self.p_drop > 0 and training and not fused:
                call('subgnn_dropout', ptr(self.dOUT[k - 1]), ptr(self.dOUT[k - 1]), M * 2 * H, self.p_drop, seed, 8 + k, step_dev, st)
        for a_ in aux:
            cur.wait_stream(a_)


    def _backward_gemms_ws(self, k, dG, dG_last, x_ptr, ldx, din, full, n_out, dE_ptr, dense, dense_dx, st, cur, aux, side=None):
        """the consumers of layer k's gate gradients on the TMA-fed grouped kernel (tcgemm_ws.cu): input gradient (the chain) and the
        three weight gradients.  grouped == 2: ONE launch per layer; grouped == 1: one launch for the LAST processed layer (nothing
        of the chain follows it: the tail of the backward pass), otherwise the input gradient on the chain stream and the weight
        gradients as one companion launch on an auxiliary stream, capped to the SMs the chain launch leaves free; grouped == 0:
        one launch per product."""
        a, H, D, M, T = self.arena, self.H, self.D, self.n_seq * self.T, self.T
        o = a.lstm_off[k]
        gd = _abi.gemm_desc
        w_ih, gw_ih, gw_hh = a.base_addr(o['weight_ih']), a.base_addr(o['weight_ih'], 'grads'), a.base_addr(o['weight_hh'], 'grads')
        for a_ in aux:                                         # every auxiliary stream joins the capture / the step's dependency chain
            a_.wait_stream(cur)
        bw = [gd(_abi.GEMM_BWD_WEIGHT, dG, 8 * H, x_ptr, ldx, gw_ih, din, M, n_out, din)]
        if not full:
            xl, ldxl, _ = self._last_rows(k, x_ptr, ldx, None)
            bw.append(gd(_abi.GEMM_BWD_WEIGHT, dG_last, T * 8 * H, xl, ldxl, gw_ih + 4 * (4 * H * din), din, self.n_seq, 4 * H, din))
        for d_ in range(2 if full else 1):                     # reverse direction took one step from h = 0: no W_hh gradient
            bw.append(gd(_abi.GEMM_BWD_WEIGHT_SHIFT, dG + 4 * (d_ * 4 * H), 8 * H, self.OUT[k].data_ptr() + 4 * (d_ * H), 2 * H,
                         gw_hh + 4 * (d_ * 4 * H * H), H, M, 4 * H, H, shift=-1 if d_ == 0 else 1, period=T))
        bi, bi_after = [], []
        if k > 0:
            dx_ptr, lddx, scat = ptr(self.dOUT[k - 1]), 2 * H, None
        elif dense:
            dx_ptr, lddx, scat = (ptr(dense_dx) if dense_dx is not None else None), D, None
        else:
            dx_ptr, lddx, scat = dE_ptr, D, ptr(self.ids_flat)
        if dx_ptr:
            bi.append(gd(_abi.GEMM_BWD_INPUT, dG, 8 * H, w_ih, din, dx_ptr, lddx, M, n_out, din, scatter_ids=scat, accumulate=1 if scat else 0))
            if not full:
                w_rev = w_ih + 4 * (4 * H * din)
                if scat:
                    bi.append(gd(_abi.GEMM_BWD_INPUT, dG_last, T * 8 * H, w_rev, din, dx_ptr, lddx, self.n_seq, 4 * H, din,
                                 scatter_ids=ptr(self.ids_last), accumulate=1))
                elif side is not None and k > 0:   # stored into the side buffer (rows t = T-1) by the same launch; the next BPTT adds it
                    bi.append(gd(_abi.GEMM_BWD_INPUT, dG_last, T * 8 * H, w_rev, din, side + 4 * ((T - 1) * lddx), T * lddx, self.n_seq,
                                 4 * H, din))
                elif self.bi_merge and k > 0 and k == self.nl - 1:   # both products add with atomics onto the zero-filled dOUT[k-1] (backward()): ONE launch on the chain
                    bi[0] = gd(_abi.GEMM_BWD_INPUT, dG, 8 * H, w_ih, din, dx_ptr, lddx, M, n_out, din, accumulate=2)
                    bi.append(gd(_abi.GEMM_BWD_INPUT, dG_last, T * 8 * H, w_rev, din, dx_ptr + 4 * ((T - 1) * lddx), T * lddx, self.n_seq,
                                 4 * H, din, accumulate=2))
                    cur.wait_stream(aux[1])              # the zero fill
                else:          # read-add-store onto rows the main product writes: after it, not beside it
                    bi_after.append(gd(_abi.GEMM_BWD_INPUT, dG_last, T * 8 * H, w_rev, din, dx_ptr + 4 * ((T - 1) * lddx), T * lddx, self.n_seq,
                                       4 * H, din, accumulate=1))
        mode = self.grouped
        if mode == 2 or (mode == 1 and (k == 0 or not bi)):
            _abi.gemm_group(bi + bw, st)
        elif mode == 1:
            tiles = ((M + 127) // 128) * ((din + 127) // 128)
            sms = _abi.lib.subgnn_device_sm_count()
            with torch.cuda.stream(aux[0]):
                _abi.gemm_group(bw, aux[0].cuda_stream, max_ctas=max(32, sms - min(sms, tiles)))
            _abi.gemm_group(bi, st)
        else:
            for i, d_ in enumerate(bw):
                s_ = aux[i % len(aux)]
                with torch.cuda.stream(s_):
                    _abi.gemm_group([d_], s_.cuda_stream)
            for d_ in bi:
                _abi.gemm_group([d_], st)
        for d_ in bi_after:
            _abi.gemm_group([d_], st)


# ------------------------------------------------------------------------------------------------------
class DpExchange:
    """Gradient exchange + optimizer over NVLink peer memory (csrc/dp.cu): symmetric gradient / parameter arenas
    (torch.distributed._symmetric_memory maps every peer's allocation into this process), reduce-scatter of the gradient shards by
    direct peer loads (or one in-switch multimem.ld_reduce), sharded Adam, all-gather of the updated parameters by direct peer stores
    (or one multimem.st), cross-GPU barriers inside the kernels (epoch flags in a symmetric array).  Replaces the NCCL all-reduce +
    two optimizer kernels of the data-parallel step."""

    @staticmethod
    def allocator(device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        try:
            symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
        except Exception:
            pass
        return lambda n: symm.empty(n, dtype=torch.float32, device=torch.device(device))

    def __init__(self, arena, world, rank, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = world, rank
        self.n = arena.size
        self.shard = _align((self.n + world - 1) // world)
        self.slots = symm.empty(max(world, 4), dtype=torch.float32, device=torch.device(device)).zero_()
        self.flags = symm.empty(int(_abi.lib.subgnn_dp_flag_words()), dtype=torch.int32, device=torch.device(device)).zero_()
        self.epoch = torch.ones(1, dtype=torch.int32, device=torch.device(device))      # advanced by the closing kernel of every exchange
        group = dist.group.WORLD
        self.h_g, self.h_p, self.h_s = symm.rendezvous(arena.grads, group), symm.rendezvous(arena.params, group), symm.rendezvous(self.slots, group)
        self.h_f = symm.rendezvous(self.flags, group)
        assert int(self.h_g.buffer_ptrs[rank]) == arena.grads.data_ptr() and int(self.h_p.buffer_ptrs[rank]) == arena.params.data_ptr()
        tab = lambda h: (C.c_ulonglong * world)(*[int(x) for x in h.buffer_ptrs])
        self.pg, self.pp, self.ps, self.pf = tab(self.h_g), tab(self.h_p), tab(self.h_s), tab(self.h_f)
        torch.cuda.synchronize()
        dist.barrier()                                 # every rank's flag array is zeroed before anyone signals into it
        self.gsum = torch.zeros(self.shard, dtype=torch.float32, device=torch.device(device))
        self.arena = arena
        self.in_kernel_barrier = os.environ.get('SUBGNN_DP_BARRIER', 'torch') == 'kernel'
        # NVLS multicast mappings (in-switch reduction / broadcast) when the fabric offers them; SUBGNN_DP_MULTICAST=0: per-peer loads / stores
        self.mc_g = self.mc_p = None
        if _flag('SUBGNN_DP_MULTICAST', world > 2):
            try:
                mg, mp = int(self.h_g.multicast_ptr or 0), int(self.h_p.multicast_ptr or 0)      # 0: the fabric / driver offers no multicast
                if mg and mp:
                    self.mc_g, self.mc_p = mg, mp
            except Exception:                         # noqa: BLE001 — no multicast: peer loads / stores
                self.mc_g = self.mc_p = None
        if rank == 0 and os.environ.get('SUBGNN_B200_VERBOSE'):
            print('[subgnn_b200] data-parallel exchange over NVLink peer memory: world %d, shard %d floats, NVLS multicast %s' %
                  (world, self.shard, 'on' if self.mc_g else 'off'), file=sys.stderr)

    def step(self, lr, step_dev, clip, st):
        a = self.arena
        # cross-GPU barriers: symmetric-memory signal-pad barriers (1-warp kernels) around the two exchange kernels, or
        # (SUBGNN_DP_BARRIER=kernel) epoch flags inside the kernels.  Measured (tools/dp_bench.py): 2 GPUs 46.6 vs 50.3 us per exchange,
        # 8 GPUs: flags polled by every block of 8 ranks are slower still (78 - 110 us) — the barrier kernels stay the default
        in_kernel = self.in_kernel_barrier
        ep = ptr(self.epoch) if in_kernel else None
        if not in_kernel:
            self.h_g.barrier(channel=0)                # every rank's gradient arena is complete
        call('subgnn_dp_reduce_scatter', self.pg, self.ps, self.pf, self.mc_g, ep, self.world, self.rank, self.n, self.shard, ptr(self.gsum), st)
        if not in_kernel:
            self.h_g.barrier(channel=1)                # shard sums of squares published; peers are done reading my gradients
        call('subgnn_dp_adam_allgather', self.pp, self.pf, self.mc_p, ep, self.world, self.rank, self.n, self.shard, ptr(self.gsum), ptr(a.m),
             ptr(a.v), lr, 0.9, 0.999, 1e-8, ptr(step_dev), ptr(self.slots), clip, 1.0 / self.world, st)
        if not in_kernel:
            self.h_g.barrier(channel=2)                # every shard of my parameter arena has been written

    def gather_moments(self):
        """full Adam moments on every rank (checkpoints): each rank holds only its own shard"""
        import torch.distributed as dist
        out = []
        for t in (self.arena.m, self.arena.v):
            full = torch.zeros(self.shard * self.world, dtype=torch.float32, device=t.device)
            lo, hi = self.rank * self.shard, min(self.n, (self.rank + 1) * self.shard)
            if hi > lo:
                full[lo:hi] = t[lo:hi]
            dist.all_reduce(full)
            out.append(full[:self.n])
        return out


# ------------------------------------------------------------------------------------------------------
class StepContext:
    """All per-step scratch + the C descriptor for one (split tables, batch size, training flag)."""

    def __init__(self, eng, tables, B, training):
        hp, a, dev = eng.hp, eng.arena, eng.device
        D, L = hp['node_embed_size'], hp['n_layers']
        self.B, self.training, self.tables = B, training, tables
        self.R_cap = B * tables.max_cc_per_sub
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        zi = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        h1, h2, K, hid = hp['linear_hidden_dim_1'], hp['linear_hidden_dim_2'], eng.num_classes, eng.hid_dim
        self.batch_idx = zi(B)
        # pinned staging ring for the step's only host input: a slot is rewritten only after the copy that read it has retired
        # (steps are enqueued asynchronously, the host may run several steps ahead of the device)
        self.batch_ring = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(8)]
        self.ring_events = [None, None]
        self.ring_pos = 0
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()     # written by a D2H copy node at the end of the step
        self.b_rowptr, self.meta = zi(B + 1), zi(4)
        self.row_b, self.row_g = zi(self.R_cap), zi(self.R_cap)
        A_pi, A_pb, A_s = eng.A['pi'], eng.A['pb'], eng.A['s']
        # gradient-side scratch that must start at zero every step lives in ONE buffer (single fill)
        n_dq = L * (B * A_pi + A_pb + 2 * A_s)
        self.zero_scratch = z(_align(n_dq) + 8 + _align(B * hid) + _align(B * h1) + 4)
        self.dq_pi = self.zero_scratch[:L * B * A_pi]
        self.dq_pb = self.zero_scratch[L * B * A_pi:L * B * A_pi + L * A_pb]
        self.dq_s = self.zero_scratch[L * (B * A_pi + A_pb):n_dq]
        self.sumsq = self.zero_scratch[_align(n_dq):_align(n_dq) + 1]
        self.q_pi, self.q_pb, self.q_s = z(max(L * B * A_pi, 1)), z(max(L * A_pb, 1)), z(max(L * 2 * A_s, 1))
        self.n_wt = z(max(L * 2 * 2 * D * D, 1))
        self.lin_wt = [z(h1 * hid), z(h2 * h1), z(K * h2)]
        self.X0 = z(self.R_cap, D)
        nN = L if hp['use_neighborhood'] else 0
        self.Nh, self.Nagg, self.Ndpre = z((nN + 1) * 2 * self.R_cap * D), z(max(nN, 1) * 2 * self.R_cap * D), z(max(nN, 1) * 2 * self.R_cap * D)
        # Z (atomic readout target) and H1 (split-K target of the first MLP layer) are zeroed per step with the scratch
        self.grad_zero = self.zero_scratch[:_align(n_dq) + 8]
        self.fwd_zero = self.zero_scratch[_align(n_dq) + 8:]
        self.Z = self.fwd_zero[:_align(B * hid)]
        self.H1 = self.fwd_zero[_align(B * hid):_align(B * hid) + B * h1].view(B, h1)
        self.H2, self.logits, self.loss_b = z(B, h2), z(B, K), z(B)
        self.dlogits, self.dH2, self.dH1, self.dZ = z(B, K), z(B, h2), z(B, h1), z(B, hid)
        self.loss = self.fwd_zero[_align(B * hid) + _align(B * h1):_align(B * hid) + _align(B * h1) + 1]   # zeroed with Z / H1: the readout kernel adds into it
        d = ModelDesc()
        t = tables
        P = lambda x: ptr(x) if x is not None else None
        d.sub_ccptr, d.sub_maxlen, d.cc_nodeptr, d.cc_nodes = P(t.sub_ccptr), P(t.sub_maxlen), P(t.cc_nodeptr), P(t.cc_nodes)
        for s in range(2):
            d.n_ids[s], d.n_sim[s], d.p_sim[s], d.s_sim[s] = P(t.n_ids[s]), P(t.n_sim[s]), P(t.p_sim[s]), P(t.s_sim[s])
        d.p_int_ids, d.p_bor_ids = P(t.p_int_ids), P(t.p_bor_ids)
        if t.multilabel:
            d.labels_multi = P(t.labels)
        else:
            d.labels = P(t.labels)
        d.E = eng.E_ptr()
        d.dE = eng.dE_ptr() if training else None
        for ci, ch in enumerate(CHANNELS):
            if ch in a.mpn_base:
                d.mpn_params[ci] = a.base_addr(a.mpn_base[ch])
                d.mpn_grads[ci] = a.base_addr(a.mpn_base[ch], 'grads') if training else None
        if a.cc_tables:
            tabs = eng.cc_tables_for(t)
            for s in range(2):
                d.cc_tab[s] = tabs[s][0]
                d.cc_tab_grad[s] = tabs[s][1] if training else None
        for i, nm in enumerate(('lin', 'lin2', 'lin3')):
            d.lin_w[i], d.lin_b[i] = a.addr(nm + '.weight'), a.addr(nm + '.bias')
            d.lin_gw[i] = a.addr(nm + '.weight', 'grads') if training else None
            d.lin_gb[i] = a.addr(nm + '.bias', 'grads') if training else None
            d.lin_wt[i] = ptr(self.lin_wt[i])
        if eng.lstm is not None:
            d.emb_s, d.d_emb_s = ptr(eng.lstm.EMB), ptr(eng.lstm.dEMB)
        d.batch_idx, d.step_dev, d.b_rowptr, d.meta = ptr(self.batch_idx), ptr(eng.step_dev), ptr(self.b_rowptr), ptr(self.meta)
        d.row_b, d.row_g = ptr(self.row_b), ptr(self.row_g)
        d.n_wt = ptr(self.n_wt)
        d.q_pi, d.q_pb, d.q_s = ptr(self.q_pi), ptr(self.q_pb), ptr(self.q_s)
        d.dq_pi, d.dq_pb, d.dq_s = (self.dq_pi.data_ptr(), self.dq_pb.data_ptr(), self.dq_s.data_ptr())
        d.X0, d.Nh, d.Nagg, d.Ndpre = ptr(self.X0), ptr(self.Nh), ptr(self.Nagg), ptr(self.Ndpre)
        d.Z, d.H1, d.H2, d.logits, d.loss_b = ptr(self.Z), ptr(self.H1), ptr(self.H2), ptr(self.logits), ptr(self.loss_b)
        d.loss_sum = self.loss.data_ptr()
        d.mlp_fused = 0
        # N-channel weight gradients: many small CTAs when the launch ends the backward pass (one-layer walk encoder), one 512-row
        # chunk per product when it hides behind a longer BPTT chain (model.cu: subgnn_model_wgrad)
        d.wgrad_rows = 64 if (eng.lstm is None or eng.lstm.nl == 1) else 0
        d.dlogits, d.dH2, d.dH1, d.dZ = ptr(self.dlogits), ptr(self.dH2), ptr(self.dH1), ptr(self.dZ)
        d.seed = eng.seed
        d.n_sub, d.n_cc, d.n_nodes = t.n_sub, t.n_cc, eng.n_nodes
        d.D, d.L, d.hid, d.h1, d.h2, d.n_classes = D, L, hid, h1, h2, K
        d.use_n, d.use_p, d.use_s = int(hp['use_neighborhood']), int(hp['use_position']), int(hp['use_structure'])
        d.A_ni, d.A_nb, d.A_pi, d.A_pb, d.A_s = eng.A['ni'], eng.A['nb'], A_pi, A_pb, A_s
        d.pool_max = 1 if hp['cc_aggregator'] == 'max' else 0
        d.trainable_cc, d.C_pad = int(a.cc_tables), t.C_pad
        d.multilabel, d.training, d.use_proj = int(t.multilabel), int(training), int(hp['use_mpn_projection'])
        d.B, d.R_cap, d.step, d.lin_dropout = B, self.R_cap, 0, float(hp['lin_dropout'])
        self.desc = d
        self.dptr = C.addressof(d)
        # readout section as one cluster kernel (model.cu readout_cluster_kernel): measured on B200 at the PPI-BP shape (same box,
        # 100 steps): 0.3372 ms/step with the three batch-level MLP kernels, 0.3365 (first-layer slices staged by threads, dZ per thread)
        # ... 0.3498 (bulk-copy staging, warp-per-column dZ) with the cluster kernel, 0.3429 when it also produces the MLP weight
        # gradients (atomics on the critical chain) — no gain: a single-shot kernel of ~10 dependent phases is bound by instruction
        # fetch and L2 round trips, not by the work.  It stays an opt-in (SUBGNN_READOUT_CLUSTER=1 / hp['b200_readout_cluster']),
        # parity-tested like the default path.
        self.readout_cluster = _flag('SUBGNN_READOUT_CLUSTER', hp.get('b200_readout_cluster', False)) and bool(_abi.lib.subgnn_model_readout_supported(self.dptr))
        self.graph = None
        self.generation = 0          # forwards run on this context (an autograd backward checks it reads its own forward's buffers)


class Engine:
    def __init__(self, hp, prepared, device='cuda', graph=None, seed=0, world_size=1):
        self.hp = dict(hp)
        hp = self.hp
        if hp.get('batch_norm', False) or hp.get('ff_attn', False) or hp.get('norm_pos_struc_embed', False):
            raise NotImplementedError('batch_norm / ff_attn / norm_pos_struc_embed are off in every shipped config and outside the fused path')
        if hp['cc_aggregator'] not in ('sum', 'max'):
            raise NotImplementedError(hp['cc_aggregator'])
        self.device = torch.device(device)
        self.prepared, self.graph = prepared, graph
        self.seed, self.world_size = int(seed), world_size
        D, L = hp['node_embed_size'], hp['n_layers']
        self.n_nodes = int(np.asarray(prepared['embeddings']).shape[0]) - 1
        self.num_classes = int(prepared['num_classes'])
        self.A = {'ni': hp['n_anchor_patches_N_in'] if hp['use_neighborhood'] else 0, 'nb': hp['n_anchor_patches_N_out'] if hp['use_neighborhood'] else 0,
                  'pi': hp['n_anchor_patches_pos_in'] if hp['use_position'] else 0, 'pb': hp['n_anchor_patches_pos_out'] if hp['use_position'] else 0,
                  's': hp['n_anchor_patches_structure'] if hp['use_structure'] else 0}
        self.hid_dim = D + L * ((2 * D if hp['use_neighborhood'] else 0) + self.A['pi'] + self.A['pb'] + 2 * self.A['s'])   # SubGNN.py:118-147
        self.tables = {}
        self.tables['train'] = SplitTables(prepared, 'train', hp, self.device, graph)
        tt = self.tables['train']
        # data parallel: gradient exchange + optimizer over NVLink peer memory (DpExchange) when symmetric memory is available;
        # SUBGNN_DP_FUSED=0 (or a failed rendezvous) keeps the NCCL all-reduce captured inside the step graph
        self.dp, alloc = None, None
        if world_size > 1 and _flag('SUBGNN_DP_FUSED', hp.get('b200_dp_fused', True)):
            try:
                alloc = DpExchange.allocator(self.device)
                alloc(16)
            except Exception as e:                    # noqa: BLE001 — any failure here means: no peer memory, use NCCL
                print('[subgnn_b200] symmetric memory unavailable (%s): data-parallel exchange through NCCL' % str(e).splitlines()[0], file=sys.stderr)
                alloc = None
        self.arena = ParamArena(hp, self.n_nodes, self.num_classes, self.hid_dim, tt.n_sub, tt.C_pad, self.device, alloc=alloc)
        if alloc is not None:
            import torch.distributed as dist
            try:
                self.dp = DpExchange(self.arena, world_size, dist.get_rank(), self.device)
            except Exception as e:                    # noqa: BLE001
                print('[subgnn_b200] symmetric-memory rendezvous failed (%s): data-parallel exchange through NCCL' % str(e).splitlines()[0], file=sys.stderr)
                self.dp = None
        emb = torch.as_tensor(np.asarray(prepared['embeddings']), dtype=torch.float32).to(self.device)
        if self.arena.embed_trainable:
            self.arena.view('node_embeddings.weight').copy_(emb)
            self.E_frozen = None
        else:
            self.E_frozen = emb.contiguous()
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lstm = None
        if hp['use_structure']:
            self.lstm = LstmRunner(self.arena, hp, self._structure_walks(prepared), 2 * L * self.A['s'], self.device)
        self.ctx = {}
        self.eval_cc = {}
        self.lr = float(hp['learning_rate'])
        self.grad_clip = float(hp.get('grad_clip', 0.0) or 0.0)
        self.concurrent = bool(hp.get('b200_concurrent_streams', True))
        self._side = None

    def init_parameters(self, seed=0):
        """Random initialisation with torch's default laws for the reference's modules: nn.Linear
        U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias, nn.LSTM U(-1/sqrt(H), 1/sqrt(H)); the embedding
        table keeps the supplied (pre-trained / synthetic) rows; trainable cc tables start from the pooling."""
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(seed))
        a = self.arena
        H = self.hp['node_embed_size']
        for name, (off, shape) in a.entries.items():
            if name == 'node_embeddings.weight' or name.endswith('_cc_embed'):
                continue
            if name.startswith('lstm.lstm.'):
                bound = 1.0 / math.sqrt(H)
            elif name.endswith('.weight'):
                bound = 1.0 / math.sqrt(shape[-1])
            else:                                           # bias: fan_in of the matching weight
                bound = 1.0 / math.sqrt(a.entries[name[:-4] + 'weight'][1][-1])
            a.view(name).copy_((torch.rand(shape, generator=gen, device=self.device) * 2 - 1) * bound)
        self.init_cc_tables_from_pooling()

    # ---- pointers --------------------------------------------------------------------------------
    def E_ptr(self):
        return self.arena.addr('node_embeddings.weight') if self.arena.embed_trainable else self.E_frozen.data_ptr()

    def dE_ptr(self):
        return self.arena.addr('node_embeddings.weight', 'grads') if self.arena.embed_trainable else None

    def cc_tables_for(self, tables):
        a = self.arena
        if tables is self.tables['train']:
            return [(a.addr('train_N_I_cc_embed'), a.addr('train_N_I_cc_embed', 'grads')),
                    (a.addr('train_N_B_cc_embed'), a.addr('train_N_B_cc_embed', 'grads'))]
        snap = self.eval_cc[id(tables)]
        return [(snap.data_ptr(), None), (snap.data_ptr(), None)]

    def tables_for(self, split):
        if split not in self.tables:
            self.tables[split] = SplitTables(self.prepared, split, self.hp, self.device, self.graph)
            if self.arena.cc_tables:
                self.snapshot_eval_cc_tables(split)
        return self.tables[split]

    def snapshot_eval_cc_tables(self, split):
        """SubGNN.py:659-668: val/test channel tables are pooled once (prepare_data time)."""
        t = self.tables[split]
        cc = torch.from_numpy(np.asarray(self.prepared['cc_ids'][split])).to(self.device)
        E = self.arena.view('node_embeddings.weight') if self.arena.embed_trainable else self.E_frozen
        e = E[cc]                                     # setup-time plumbing, not on the step path
        self.eval_cc[id(t)] = (e.sum(dim=2) if self.hp['cc_aggregator'] == 'sum' else e.max(dim=2)[0]).contiguous()

    def init_cc_tables_from_pooling(self):
        """SubGNN.py:629-635: the six trainable tables start as the pooled component embeddings."""
        if not self.arena.cc_tables:
            return
        cc = torch.from_numpy(np.asarray(self.prepared['cc_ids']['train'])).to(self.device)
        E = self.arena.view('node_embeddings.weight') if self.arena.embed_trainable else self.E_frozen
        e = E[cc]
        pooled = e.sum(dim=2) if self.hp['cc_aggregator'] == 'sum' else e.max(dim=2)[0]
        for nm in ('N_I', 'N_B', 'S_I', 'S_B', 'P_I', 'P_B'):
            self.arena.view('train_%s_cc_embed' % nm).copy_(pooled)

    def add_split(self, split, q):
        """merge the products of a later-prepared split (prepare_test_data, SubGNN.py:994-1022) into the engine."""
        p = self.prepared
        for key in ('cc_ids', 'labels', 'sub_G', 'N_border', 'I_S_sim', 'B_S_sim', 'anchors_neigh_int', 'anchors_neigh_border', 'anchors_pos_int'):
            if isinstance(q.get(key), dict) and split in q[key]:
                if not isinstance(p.get(key), dict):
                    p[key] = {}
                p[key][split] = q[key][split]
        if isinstance(q.get('NP_sim'), dict) and split in q['NP_sim'] and isinstance(p.get('NP_sim'), dict):
            p['NP_sim'][split] = q['NP_sim'][split]
        elif isinstance(p.get('NP_sim'), dict) and self.graph is not None and self.graph.hop is not None:
            pass                                              # SplitTables falls back to the hop table for this split
        self.tables.pop(split, None)

    def rebind(self, prepared):
        """swap in freshly sampled anchors (resample_anchor_patches, SubGNN.py:449-457); parameters and Adam state stay."""
        hp = self.hp
        self.prepared = prepared
        self.tables = {'train': SplitTables(prepared, 'train', hp, self.device, self.graph)}
        self.ctx, self.eval_cc = {}, {}
        if hp['use_structure']:
            self.lstm = LstmRunner(self.arena, hp, self._structure_walks(prepared), 2 * hp['n_layers'] * self.A['s'], self.device)

    def _structure_walks(self, prepared):
        T = self.hp['random_walk_len']
        walks = []
        for key in (2, 3):                                           # internal walks, then border walks (side-major)
            for l in range(self.hp['n_layers']):
                walks.append(np.asarray(prepared['anchors_structure'][l][key]).reshape(-1, T))
        return torch.from_numpy(np.concatenate(walks).astype(np.int32)).to(self.device)

    def context(self, split, B, training):
        key = (split, B, training)
        if key not in self.ctx:
            self.ctx[key] = StepContext(self, self.tables_for(split), B, training)
        return self.ctx[key]

    # ---- launches --------------------------------------------------------------------------------
    # The walk-encoder LSTM depends only on the parameters (the structure anchor patches are shared by every subgraph,
    # anchor_patch_samplers.py:381-386), the neighbourhood chains only on the batch: they are launched on two streams
    # (fork / join with events, captured into the step graph as parallel branches).
    def _side_stream(self):
        # optional high priority for the LSTM chain (the critical path of the step).  Measured on B200 (tools/ab_bench.sh, PPI-BP
        # shape, same box): no effect (0.4362 -> 0.4366 ms/step), so it is off by default; kept as a tuning switch
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1 if _flag('SUBGNN_CHAIN_PRIORITY', self.hp.get('b200_chain_priority', False)) else 0)
        return self._side

    def _prep_stream(self):
        if getattr(self, '_prep', None) is None:
            self._prep = torch.cuda.Stream(device=self.device)
        return self._prep

    def _q_stream(self):
        if getattr(self, '_qs', None) is None:
            self._qs = torch.cuda.Stream(device=self.device)
        return self._qs

    def _forward_launches(self, c, st, zero_grads=False, split=False, inc_step=False):
        """split (fused training step only): only the structure-channel columns of Z depend on the LSTM, so the position outputs and
        the first MLP layer over every other column run beside the LSTM chain; behind the join stay the structure outputs, their
        slices of the first layer, the rest of the MLP and the structure columns of dZ (the others are produced in the backward
        pass, beside the BPTT chain)."""
        main = torch.cuda.current_stream()
        fork = self.lstm is not None and self.concurrent
        # launch order of the two branches inside the captured graph: [r2] with the faster projection GEMMs the neighbourhood branch
        # (zero fills -> batch prep -> weight transposes -> position q -> row_fwd phase 1, ~75 us of dependent launches, starved
        # while the first recurrence owns every SM) ends AFTER the LSTM chain; its nodes are recorded first (measured, same box:
        # 0.3182 -> 0.3140 ms/step; SUBGNN_NBRANCH_FIRST=0 restores the round-1 order)
        nfirst = fork and _flag('SUBGNN_NBRANCH_FIRST', self.hp.get('b200_nbranch_first', True))
        step_ev = None
        # fork_early: the LSTM chain forks before the step-counter kernel and waits for it (event) in front of the first recurrence.
        # Measured on B200 (PPI-BP shape, same box, 100 steps): 0.3225 ms/step with, 0.3101 without — the second incoming edge
        # takes the programmatic (PDL) edge projection -> recurrence away — so it is off.
        fork_early = _flag('SUBGNN_FORK_EARLY', self.hp.get('b200_fork_early', False))
        if inc_step and not (fork and fork_early):
            call('subgnn_inc_step', ptr(self.step_dev), st)
            inc_step = False
        if fork:                                  # the LSTM chain is the longest of the step: it starts before the zero fills
            side = self._side_stream()
            side.wait_stream(main)                # fork_early: BEFORE the step-counter kernel (gather + first projection do not read it)
            self._prep_stream().wait_stream(main)
        if inc_step:
            call('subgnn_inc_step', ptr(self.step_dev), st)
            step_ev = torch.cuda.Event()
            step_ev.record(main)
        if fork and not nfirst:
            with torch.cuda.stream(side):
                self.lstm.forward(self.E_ptr(), c.training, self.seed, ptr(self.step_dev), side.cuda_stream, step_event=step_ev)
        if zero_grads:
            self.zero_grads(c, st, include_fwd=True)
        else:
            call('subgnn_fill_zero', ptr(c.fwd_zero), c.fwd_zero.numel(), st)
        if self.lstm is not None and not fork:
            self.lstm.forward(self.E_ptr(), c.training, self.seed, ptr(self.step_dev), st)
        # Neighbourhood branch.  Its long kernel (row_fwd phase 1: pooling + N chains) needs the transposed weights and the batch
        # prep, NOT the position q (only phase 2 reads q).  SUBGNN_FWD_BRANCHES=1 starts the transposes with the step on their own
        # branch and runs the position q beside row_fwd instead of before it.  Measured (tools/ab_bench.sh, PPI-BP shape, same
        # box): 0.3751 -> 0.3831 ms/step — released earlier, row_fwd's CTAs land on the SMs together with the layer-1 projection
        # of the LSTM chain (the critical path) instead of beside the half-empty top-layer recurrence — so it stays a switch.
        bfork = fork and _flag('SUBGNN_FWD_BRANCHES', False)
        pfork = bfork or (fork and _flag('SUBGNN_PREP_BRANCH', False))
        if pfork:
            prep = self._prep_stream()
            with torch.cuda.stream(prep):
                call('subgnn_model_prep_weights', c.dptr, prep.cuda_stream)
        call('subgnn_model_prep_batch', c.dptr, st)
        if not pfork:
            call('subgnn_model_prep_weights', c.dptr, st)
        if bfork:
            qs = self._q_stream()
            qs.wait_stream(main)
            with torch.cuda.stream(qs):
                call('subgnn_model_q_fwd_part', c.dptr, 1, qs.cuda_stream)   # position anchors
        else:
            call('subgnn_model_q_fwd_part', c.dptr, 1, st)           # position anchors
        if pfork:
            main.wait_stream(prep)
        call('subgnn_model_rows_fwd', c.dptr, 1, st)                 # pooling + neighbourhood channel
        if nfirst:
            with torch.cuda.stream(side):
                self.lstm.forward(self.E_ptr(), c.training, self.seed, ptr(self.step_dev), side.cuda_stream, step_event=step_ev)
        if bfork:
            main.wait_stream(qs)
        if split:
            call('subgnn_model_rows_fwd', c.dptr, 2, st)             # position property-aware outputs
            call('subgnn_model_mlp_stage', c.dptr, 1, 1, st)         # first MLP layer over the LSTM-independent columns of Z
            main.wait_stream(side)
            call('subgnn_model_q_fwd_part', c.dptr, 2, st)           # structure anchors (LSTM output)
            call('subgnn_model_rows_fwd', c.dptr, 4, st)             # structure property-aware outputs
            call('subgnn_model_mlp_stage', c.dptr, 1, 2, st)         # their slices of the first layer
            call('subgnn_model_mlp_stage', c.dptr, 2, 0, st)         # lin2, lin3, loss, d logits, dH2, dH1
            if split == 'fwd':                                       # forward half only: all of dZ here, the usual backward pass behind it
                call('subgnn_model_mlp_stage', c.dptr, 4, 0, st)
                call('subgnn_sum_to_scalar', ptr(c.loss_b), c.B, ptr(c.loss), st)
            else:
                call('subgnn_model_mlp_stage', c.dptr, 4, 2, st)     # structure columns of dZ (the chain continues through them)
            return
        if fork:
            main.wait_stream(side)
        call('subgnn_model_q_fwd_part', c.dptr, 2, st)               # structure anchors (LSTM output)
        call('subgnn_model_rows_fwd', c.dptr, 6, st)                 # P / S property-aware outputs
        if c.readout_cluster:
            call('subgnn_model_readout', c.dptr, st)                 # lin -> lin2 -> lin3 -> loss (-> dZ [, MLP gradients]) in one launch
        else:
            call('subgnn_model_mlp_fwd', c.dptr, st)
            call('subgnn_sum_to_scalar', ptr(c.loss_b), c.B, ptr(c.loss), st)

    def _backward_launches(self, c, st, external_dlogits=False, split=False):
        main = torch.cuda.current_stream()
        if split and split != 'fwd':
            call('subgnn_model_rows_bwd', c.dptr, 4, st)             # structure outputs: d q_s, d b_p
            call('subgnn_model_q_bwd_part', c.dptr, 2, st)           # d emb_s
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.lstm.backward(self.E_ptr(), self.dE_ptr(), c.training, self.seed, ptr(self.step_dev), side.cuda_stream)
            call('subgnn_model_mlp_stage', c.dptr, 4, 1, st)         # every other column of dZ
            call('subgnn_sum_to_scalar', ptr(c.loss_b), c.B, ptr(c.loss), st)
            call('subgnn_model_rows_bwd', c.dptr, 2, st)             # position outputs: d q_p, d b_p
            call('subgnn_model_q_bwd_part', c.dptr, 1, st)
            call('subgnn_model_rows_bwd', c.dptr, 1, st)             # neighbourhood chains + pooling
            call('subgnn_model_wgrad', c.dptr, st)
            main.wait_stream(side)
            return
        if external_dlogits:
            call('subgnn_model_mlp_bwd', c.dptr, st)
        call('subgnn_model_rows_bwd', c.dptr, 6, st)                 # d q, d b_p
        fork = self.lstm is not None and self.concurrent
        split_q = fork and _flag('SUBGNN_Q_BWD_SPLIT', True)
        # the structure groups of q_bwd produce d emb_s, the input of the LSTM head gradient: on the chain; the position groups
        # (B * A_pi + A_pb rows scattered into dE) run beside the BPTT chain
        call('subgnn_model_q_bwd_part', c.dptr, 2 if split_q else 3, st)
        if fork:
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.lstm.backward(self.E_ptr(), self.dE_ptr(), c.training, self.seed, ptr(self.step_dev), side.cuda_stream)
        elif self.lstm is not None:
            self.lstm.backward(self.E_ptr(), self.dE_ptr(), c.training, self.seed, ptr(self.step_dev), st)
        if split_q:
            call('subgnn_model_q_bwd_part', c.dptr, 1, st)
        call('subgnn_model_rows_bwd', c.dptr, 1, st)                 # neighbourhood chains + pooling
        call('subgnn_model_wgrad', c.dptr, st)
        if fork:
            main.wait_stream(side)

    def zero_grads(self, c, st, include_fwd=False):
        """gradient arena + gradient-side scratch; include_fwd also clears the forward accumulation targets (Z, H1) that share
        the scratch buffer — only valid BEFORE the forward pass (the backward pass reads both)."""
        call('subgnn_fill_zero', ptr(self.arena.grads), self.arena.size, st)
        z = c.zero_scratch if include_fwd else c.grad_zero
        call('subgnn_fill_zero', ptr(z), z.numel(), st)

    def _optimizer_launches(self, c, st):
        a = self.arena
        clip = self.grad_clip
        if _flag('SUBGNN_FUSED_ADAM', self.hp.get('b200_fused_adam', False)):      # norm + clip + Adam in one launch (optim.cu)
            call('subgnn_clip_adam_step', ptr(a.params), ptr(a.grads), ptr(a.m), ptr(a.v), a.size, self.lr, 0.9, 0.999, 1e-8,
                 ptr(self.step_dev), c.sumsq.data_ptr(), clip, 1.0 / self.world_size, st)
            return
        call('subgnn_grad_sumsq', ptr(a.grads), a.size, c.sumsq.data_ptr(), st)
        call('subgnn_adam_step', ptr(a.params), ptr(a.grads), ptr(a.m), ptr(a.v), a.size, self.lr, 0.9, 0.999, 1e-8, ptr(self.step_dev),
             c.sumsq.data_ptr(), clip, 1.0 / self.world_size, st)

    def set_batch(self, c, indices):
        """host -> device copy of the step's only input: the subgraph indices (numpy array, list or CPU tensor, any int type).
        Staged through a ring of pinned buffers guarded by two events (one per half ring), so that the host, which may run
        several steps ahead of the device, never rewrites a buffer whose copy is still in flight."""
        ring = c.batch_ring
        pos, half = c.ring_pos, len(ring) // 2
        c.ring_pos = (pos + 1) % len(ring)
        if pos % half == 0:                       # entering a half ring: the copies issued from it one lap ago must have retired
            ev = c.ring_events[pos // half]
            if ev is not None:
                ev.synchronize()
        buf = ring[pos]
        if isinstance(indices, torch.Tensor):
            buf.copy_(indices.view(-1))           # converts int64 -> int32 on the fly; raises on a size mismatch
        else:
            idx = np.asarray(indices)
            assert idx.size == c.B
            buf.copy_(torch.from_numpy(idx.reshape(-1)))
        c.batch_idx.copy_(buf, non_blocking=True)
        if pos % half == half - 1:                # leaving a half ring
            if c.ring_events[pos // half] is None:
                c.ring_events[pos // half] = torch.cuda.Event()
            c.ring_events[pos // half].record()

    # ---- public API --------------------------------------------------------------------------------
    def forward(self, split, indices, training=False):
        """logits (B, K) for the subgraphs ``indices`` of ``split`` (SubGNN.forward)."""
        c = self.context(split, len(indices), training)
        self.set_batch(c, indices)
        st = _abi.stream_ptr()
        if training:
            # a training forward outside the fused step (SubGNN.training_step -> autograd): the dropout masks are keyed on
            # (seed, step counter, element), so the counter advances here as it does in _grad_launches; the matching backward
            # (same counter value) must run before the next training forward — guarded by the context generation
            call('subgnn_inc_step', ptr(self.step_dev), st)
        c.generation += 1
        self._forward_launches(c, st)
        return c.logits, c.loss

    def backward(self, split, B, training=True, external_dlogits=None):
        c = self.context(split, B, training)
        st = _abi.stream_ptr()
        if external_dlogits is not None:
            c.dlogits.copy_(external_dlogits)
        self.zero_grads(c, st)
        self._backward_launches(c, st, external_dlogits is not None)

    def allreduce_grads(self):
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.arena.grads)

    def _exchange_and_optimize(self, c, st):
        """[gradient exchange,] clip, Adam: fused over NVLink peer memory when available, else NCCL all-reduce + the two optimizer kernels"""
        if self.dp is not None:
            self.dp.step(self.lr, self.step_dev, self.grad_clip, st)
        else:
            self.allreduce_grads()
            self._optimizer_launches(c, st)

    def _capture_step(self, c):
        """ONE CUDA graph for the whole step, data parallel included: the NCCL all-reduce of the gradient arena is captured between
        the backward pass and the optimizer (NCCL collectives are stream-capturable), so a replay is a single launch on every rank
        and the exchange is ordered by graph edges instead of by the host.  SUBGNN_DP_SPLIT_GRAPH=1 keeps the round-1 form (two
        graphs around an eagerly enqueued all-reduce) for A/B runs."""
        split = self.world_size > 1 and self.dp is None and _flag('SUBGNN_DP_SPLIT_GRAPH', False)
        if split:
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                self._grad_launches(c, _abi.stream_ptr())
            with torch.cuda.graph(g2):
                self._optimizer_launches(c, _abi.stream_ptr())
                c.loss_host.copy_(c.loss, non_blocking=True)
            return (g1, g2)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._grad_launches(c, _abi.stream_ptr())
            self._exchange_and_optimize(c, _abi.stream_ptr())
            c.loss_host.copy_(c.loss, non_blocking=True)
        return (g,)

    def train_step(self, indices, use_graph=False):
        """one optimisation step on the train split: forward, loss, backward, [allreduce,] clip, Adam (in place).
        With use_graph the launches are captured once and replayed as ONE CUDA graph (the NCCL all-reduce included when data
        parallel).  The graph ends with a D2H copy node of the loss into pinned host memory (read it with ``loss_value``);
        per-step host work is one pinned copy of the indices + the graph launch."""
        c = self._last_ctx = self.context('train', indices.numel() if isinstance(indices, torch.Tensor) else len(indices), True)
        self.set_batch(c, indices)
        st = _abi.stream_ptr()
        if not use_graph:
            self._grad_launches(c, st)
            self._exchange_and_optimize(c, st)
            c.loss_host.copy_(c.loss, non_blocking=True)
            return c.loss
        if c.graph is None:
            self._grad_launches(c, st)                         # warm-up (sets function attributes, NCCL communicator) — a real step
            self._exchange_and_optimize(c, st)
            c.loss_host.copy_(c.loss, non_blocking=True)
            torch.cuda.synchronize()
            n0 = _abi.lib.subgnn_launch_count()
            c.graph = self._capture_step(c)
            self.launches_per_step = int(_abi.lib.subgnn_launch_count() - n0)
            return c.loss
        if len(c.graph) == 2:
            c.graph[0].replay()
            self.allreduce_grads()
            c.graph[1].replay()
        else:
            c.graph[0].replay()
        return c.loss

    # ---- optimizer state (checkpoints) -----------------------------------------------------------------
    def optimizer_state_dict(self):
        """Adam state of the fused step in torch.optim.Adam's own state_dict layout over the arena entries in registration order
        (== SubGNN.parameters() order), so a checkpoint written in fused mode restores into either optimizer."""
        t = int(self.step_dev.item())
        state = {}
        if self.dp is not None:                          # sharded optimizer state: every rank assembles the full moments
            m_full, v_full = self.dp.gather_moments()
            mv = lambda name, which: (m_full if which == 'm' else v_full)[self.arena.entries[name][0]:self.arena.entries[name][0] + int(np.prod(self.arena.entries[name][1]))].view(self.arena.entries[name][1])
        else:
            mv = lambda name, which: self.arena.view(name, which)
        for i, name in enumerate(self.arena.entries):
            state[i] = {'step': torch.tensor(float(t)), 'exp_avg': mv(name, 'm').detach().cpu().clone(),
                        'exp_avg_sq': mv(name, 'v').detach().cpu().clone()}
        group = {'lr': self.lr, 'betas': (0.9, 0.999), 'eps': 1e-8, 'weight_decay': 0, 'amsgrad': False, 'params': list(range(len(self.arena.entries)))}
        return {'state': state, 'param_groups': [group], 'subgnn_b200_step': t, 'names': list(self.arena.entries)}

    def load_optimizer_state_dict(self, sd):
        names = sd.get('names') or list(self.arena.entries)
        t = sd.get('subgnn_b200_step')
        for i, name in enumerate(names):
            st = sd['state'].get(i)
            if st is None or name not in self.arena.entries:
                continue
            self.arena.view(name, 'm').copy_(torch.as_tensor(st['exp_avg']).to(self.device))
            self.arena.view(name, 'v').copy_(torch.as_tensor(st['exp_avg_sq']).to(self.device))
            if t is None:
                t = int(float(st['step']))
        self.step_dev.fill_(int(t or 0))
        if sd.get('param_groups'):
            self.lr = float(sd['param_groups'][0].get('lr', self.lr))
            for c in self.ctx.values():
                c.graph = None                     # the learning rate is a launch argument baked into a captured step

    def loss_value(self, B=None):
        """loss of the last enqueued train step as a Python float: waits for the stream, reads the pinned copy."""
        c = self.context('train', B, True) if B is not None else self._last_ctx
        torch.cuda.current_stream().synchronize()
        return c.loss_host.item()

    def _grad_launches(self, c, st):
        # fused step: the readout kernel's own d logits are THE gradient, so it also produces the MLP weight / bias gradients
        # (the descriptor is copied by value at every launch: the flag is per call)
        c.desc.mlp_fused = 1 if (c.readout_cluster and _flag('SUBGNN_READOUT_FUSED_WGRAD', self.hp.get('b200_readout_fused_wgrad', False))) else 0
        # split readout schedule (the LSTM-independent columns of the first MLP layer / of dZ beside the LSTM chains): measured on
        # B200, PPI-BP shape, same box: 0.3737 ms/step against 0.3203 without — the extra launches on the main stream (row kernels
        # over all rows for a few entries each, first-layer slices) delay the neighbourhood backward and its weight gradients past the
        # end of the BPTT chain and compete with the recurrences for SMs.  Kept as a switch, off.
        split = (self.lstm is not None and self.concurrent and not c.readout_cluster and
                 _flag('SUBGNN_READOUT_SPLIT', self.hp.get('b200_readout_split', False)))
        # forward half of that schedule alone (the main stream idles ~35 us before the join: position outputs and the first MLP
        # layer over the LSTM-independent columns fit there; behind the join only the structure columns' slices remain)
        if not split and self.lstm is not None and self.concurrent and not c.readout_cluster and \
                _flag('SUBGNN_READOUT_SPLIT_FWD', self.hp.get('b200_readout_split_fwd', False)):
            split = 'fwd'
        # MLP weight gradients launched right behind the readout section on a branch of their own (model.cu: subgnn_model_mlp_wgrad)
        # instead of at the end of the main stream.  Measured (same box, 100 steps): with a ONE-layer walk encoder the backward chain
        # behind the readout is short and these three kernels end the step — density 0.1741 -> 0.1464 ms/step, EM-USER 0.5709 ->
        # 0.5599; with two layers they hide behind the BPTT chain and the early launch only adds work beside its head — PPI-BP
        # 0.3189 -> 0.3225, cutratio 0.3323 -> 0.3346, HPO-METAB 0.5901 -> 0.5882.  Default: by layer count.
        one_layer = self.lstm is not None and self.lstm.nl == 1
        early = (self.concurrent and not c.desc.mlp_fused and _flag('SUBGNN_MLP_WGRAD_EARLY', self.hp.get('b200_mlp_wgrad_early', one_layer)))
        try:
            self._forward_launches(c, st, zero_grads=True, split=split, inc_step=True)
            if early:
                main, aux = torch.cuda.current_stream(), self._prep_stream()
                aux.wait_stream(main)
                with torch.cuda.stream(aux):
                    call('subgnn_model_mlp_wgrad', c.dptr, aux.cuda_stream)
                c.desc.mlp_fused = 1                 # subgnn_model_wgrad (end of the main stream) leaves the MLP products out
            self._backward_launches(c, st, split=split)
            if early:
                main.wait_stream(aux)
        finally:
            c.desc.mlp_fused = 0

    def _step_launches(self, c, st):
        self._grad_launches(c, st)
        self._exchange_and_optimize(c, st)
