"""SubGNN module with the reference's class / method / state_dict names (SubGNN/SubGNN.py:90-1164), backed by
the CUDA step engine.

Drop-in surface kept from the reference LightningModule (pytorch-lightning 0.7.1 protocol — the module is a
plain nn.Module so it does not need Lightning to be installed):
    SubGNN(hparams, graph_path, subgraph_path, embedding_path, similarities_path, shortest_paths_path,
           degree_dict_path, ego_graph_path)            SubGNN.py:94-96
    prepare_data() / prepare_test_data()                 :1024 / :994     (GPU kernels instead of Python loops)
    forward(dataset_type, ..., subgraph_idx, ...)        :225
    training_step / validation_step / test_step          :317 / :393 / :399
    configure_optimizers / backward                      :1156 / :1163
    train_dataloader / val_dataloader / test_dataloader  :1116-1151   (batches carry indices + labels only)
plus ``training_step_fused`` — the whole optimisation step (forward, loss, backward, clip, Adam) as one captured
CUDA graph, which is what bench.py measures.
"""
import math
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import PAD_VALUE, _abi
from ._abi import call, ptr
from .engine import Engine
from .graph import DeviceGraph

PROJECT_ROOT = Path('.')      # config.py:5 equivalent; set subgnn_b200.SubGNN.PROJECT_ROOT like config.PROJECT_ROOT


from .formats import SimilarityCache, load_embeddings, load_hop_table, read_edge_list, read_subgraphs  # noqa: E402,F401


class _LSTMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, keep, x, *params):
        runner, arena = module._acquire(x.shape[0], x.shape[1])
        module._pack(arena, params)
        xc = x.contiguous().float()
        st = _abi.stream_ptr()
        call('subgnn_inc_step', ptr(module._step), st)
        runner.step_dev.copy_(module._step)                   # this call's dropout counter, read again by ITS backward
        runner.forward(None, module.training, module._seed, ptr(runner.step_dev), st, dense_x=xc.view(-1, x.shape[2]))
        out = runner.EMB.clone()
        if keep:
            ctx.module, ctx.runner, ctx.arena, ctx.training = module, runner, arena, module.training
            ctx.x_shape = x.shape
        else:                                                 # no backward will come (no_grad / nothing requires grad)
            module._release(runner, arena)
        return out

    @staticmethod
    def backward(ctx, dy):
        m, runner, arena = ctx.module, ctx.runner, ctx.arena
        st = _abi.stream_ptr()
        call('subgnn_fill_zero', ptr(arena.grads), arena.size, st)
        runner.dEMB.copy_(dy.contiguous())
        dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dy.device)
        runner.backward(None, None, ctx.training, m._seed, ptr(runner.step_dev), st, dense_dx=dx.view(-1, ctx.x_shape[2]))
        grads = tuple(arena.view(n, 'grads').clone() for n in m._arena_names)
        m._release(runner, arena)
        return (None, None, dx) + grads


class LSTM(nn.Module):
    """SubGNN.py:60-88 — bidirectional LSTM walk encoder with a linear head, same constructor and parameter names
    (``lstm.weight_ih_l0`` ... ``linear.weight``); the computation runs on csrc/lstm.cu + csrc/gemm.cu."""

    def __init__(self, n_features, h, dropout=0.0, num_layers=1, batch_first=True, aggregator='last'):
        super().__init__()
        if h != n_features:
            raise NotImplementedError('the walk encoder is built for h == n_features (SubGNN.py:175)')
        if aggregator not in ('last', 'sum'):
            raise NotImplementedError                                                              # SubGNN.py:86-87
        self.num_layers, self.aggregator, self.dropout, self.h = num_layers, aggregator, dropout, h
        self.lstm = nn.LSTM(n_features, h, num_layers=num_layers, batch_first=batch_first, dropout=dropout, bidirectional=True)   # parameters only
        self.linear = nn.Linear(h * 2, n_features)
        self._runners = {}
        self._seed = 0x5eed
        self._step = None
        self._arena_names = []
        for k in range(num_layers):
            for nm in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'):
                self._arena_names += ['lstm.lstm.%s_l%d' % (nm, k), 'lstm.lstm.%s_l%d_reverse' % (nm, k)]
        self._arena_names += ['lstm.linear.weight', 'lstm.linear.bias']

    def _module_params(self):
        return [self.get_parameter(n[len('lstm.'):]) for n in self._arena_names]

    def _acquire(self, n_seq, T):
        """one runner per OUTSTANDING forward: its buffers hold the activations saved for that call's backward (the reference
        calls the shared LSTM 2 * n_layers times per step — internal / border side of every layer, anchor_patch_samplers.py:429 —
        before loss.backward()).  Runners come from a free list per (n_seq, T) and return to it in backward."""
        from .engine import LstmRunner, ParamArena
        free = self._runners.setdefault((n_seq, T), [])
        if free:
            return free.pop()
        dev = self.linear.weight.device
        hp = {'node_embed_size': self.h, 'n_layers': 1, 'freeze_node_embeds': True, 'use_neighborhood': False, 'use_position': False,
              'use_structure': False, 'lstm_n_layers': self.num_layers, 'linear_hidden_dim_1': 1, 'linear_hidden_dim_2': 1, 'trainable_cc': False,
              'n_triangular_walks': 1, 'lstm_aggregator': self.aggregator, 'lstm_dropout': self.dropout}
        arena = ParamArena(hp, 0, 1, 1, device=dev)
        runner = LstmRunner(arena, hp, None, n_seq, dev, n_seq=n_seq, T=T)
        runner.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        if self._step is None:
            self._step = torch.zeros(1, dtype=torch.int32, device=dev)
        return runner, arena

    def _release(self, runner, arena):
        self._runners[(runner.n_seq, runner.T)].append((runner, arena))

    def _pack(self, arena, params):
        for name, p in zip(self._arena_names, params):
            arena.view(name).copy_(p.detach())

    def forward(self, input):
        params = self._module_params()
        keep = torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in params))
        return _LSTMFunction.apply(self, keep, input, *params)


class _EngineFunction(torch.autograd.Function):
    """logits = engine(split, indices); backward feeds d logits to the CUDA backward and returns the parameter
    gradients (views of the gradient arena are cloned out so torch can accumulate them)."""

    @staticmethod
    def forward(ctx, module, split, indices, training, *params):
        eng = module.engine
        logits, _ = eng.forward(split, indices, training=training)
        ctx.module, ctx.split, ctx.B, ctx.training = module, split, len(indices), training
        ctx.generation = eng.context(split, len(indices), training).generation
        return logits.clone()

    @staticmethod
    def backward(ctx, dlogits):
        m = ctx.module
        eng = m.engine
        c = eng.context(ctx.split, ctx.B, ctx.training)
        if c.generation != ctx.generation:
            raise RuntimeError('SubGNN backward after another forward of the same (split, batch size, mode): the step buffers hold '
                               'the later forward\'s activations (call backward before the next forward)')
        eng.backward(ctx.split, ctx.B, training=ctx.training, external_dlogits=dlogits.contiguous())
        grads = [eng.arena.view(name, 'grads').clone() for name in m._param_names]
        return (None, None, None, None) + tuple(grads)


class SubGNN(nn.Module):
    def __init__(self, hparams, graph_path=None, subgraph_path=None, embedding_path=None, similarities_path=None,
                 shortest_paths_path=None, degree_dict_path=None, ego_graph_path=None, device=None):
        super().__init__()
        self.device_ = torch.device(device or ('cuda' if torch.cuda.is_available() else 'cpu'))
        if self.device_.type != 'cuda':
            raise RuntimeError('subgnn_b200.SubGNN needs a CUDA device: there is no CPU path')
        self.hparams = hparams
        self.graph_path, self.subgraph_path, self.embedding_path = graph_path, subgraph_path, embedding_path
        self.similarities_path, self.shortest_paths_path = similarities_path, shortest_paths_path
        self.degree_dict_path, self.ego_graph_path = degree_dict_path, ego_graph_path
        if 'structure_similarity_fn' not in self.hparams:
            self.hparams['structure_similarity_fn'] = 'dtw'                                    # SubGNN.py:186-187
        self.engine = None
        self.metric_scores = []
        self.test_results = None
        self._param_names = []
        if graph_path is not None:
            self.read_data()
            # parameters must exist once the constructor returns (train.py:307-316 restores a checkpoint before fit), so the
            # one-off preparation runs here; Trainer.fit's prepare_data() call is then a no-op
            self.prepare_data()

    # ---- construction helpers ----------------------------------------------------------------------
    @classmethod
    def from_engine(cls, engine):
        self = cls(dict(engine.hp), device=engine.device)
        self._attach(engine)
        return self

    @classmethod
    def from_prepared(cls, hparams, prepared, graph=None, device='cuda', seed=0, init_seed=None):
        self = cls(dict(hparams), device=device)
        eng = Engine(self.hparams, prepared, device=device, graph=graph, seed=seed)
        if init_seed is not None:
            eng.init_parameters(init_seed)
        self._attach(eng)
        return self

    def _attach(self, engine):
        """expose the arena as nn.Parameters under the reference's names (state_dict compatible)."""
        self.engine = engine
        self.num_classes = engine.num_classes
        self.multilabel = engine.tables['train'].multilabel
        self._param_names = list(engine.arena.entries)
        for name in self._param_names:
            p = nn.Parameter(engine.arena.view(name), requires_grad=True)
            self._register_by_path(name, p)
        if not engine.arena.embed_trainable:
            # the reference's state_dict carries the frozen table too (nn.Embedding.from_pretrained(freeze=True), SubGNN.py:568):
            # checkpoints stay loadable on either side; the tensor is the engine's own table (load_state_dict writes through)
            self._register_by_path('node_embeddings.weight', nn.Parameter(engine.E_frozen, requires_grad=False))

    def _register_by_path(self, path, param):
        mod = self
        parts = path.split('.')
        for part in parts[:-1]:
            if not hasattr(mod, part):
                setattr(mod, part, nn.Module())
            mod = getattr(mod, part)
        mod.register_parameter(parts[-1], param)

    # ---- data (SubGNN.py:519-570) --------------------------------------------------------------------
    def read_data(self):
        root = PROJECT_ROOT
        edges = read_edge_list(root / self.graph_path)
        emb = load_embeddings(root / self.embedding_path)
        n_nodes = max(int(edges.max()) + 1 if edges.size else 0, emb.shape[0])
        self.graph = DeviceGraph.from_edges(n_nodes, edges, device=self.device_, one_indexed=False)    # ids become 1-indexed (:555-559)
        first = np.full(n_nodes, n_nodes, dtype=np.int64)                                      # nx.read_edgelist node insertion order
        np.minimum.at(first, edges.reshape(-1), np.arange(edges.size))
        self.graph.insertion_rank = np.argsort(np.argsort(first, kind='stable'), kind='stable')
        splits, self.multilabel, n_labels = read_subgraphs(root / self.subgraph_path)
        self.sub_G = {k: [[n + 1 for n in s] for s in v[0]] for k, v in splits.items()}
        if self.hparams.get('subset_data', False):                                             # :542-546
            B = self.hparams['batch_size']
            self.sub_G = {k: v[:B] for k, v in self.sub_G.items()}
            splits = {k: (v[0][:B], v[1][:B]) for k, v in splits.items()}
        if self.multilabel:                                                                    # :533-536 MultiLabelBinarizer().fit(all)
            classes = sorted({l for v in splits.values() for labs in v[1] for l in labs})
            col = {c: i for i, c in enumerate(classes)}
            self.multilabel_classes = classes
            self.sub_G_label = {}
            for k, v in splits.items():
                ind = np.zeros((len(v[1]), len(classes)), dtype=np.int64)
                for i, labs in enumerate(v[1]):
                    ind[i, [col[l] for l in labs]] = 1
                self.sub_G_label[k] = ind
            self.num_classes = max(classes) + 1                                                # :548-549
            if self.num_classes != len(classes):
                raise NotImplementedError('multi-label task whose label ids are not all present in train/val/test')
        else:
            self.sub_G_label = {k: np.array([l[0] for l in v[1]], dtype=np.int64) for k, v in splits.items()}
            self.num_classes = int(max(v.max() for v in self.sub_G_label.values() if len(v))) + 1   # :551
        self.hparams['node_embed_size'] = emb.shape[1]
        self.embeddings = np.concatenate([np.zeros((1, emb.shape[1]), dtype=np.float32), emb,
                                          np.zeros((n_nodes - emb.shape[0], emb.shape[1]), dtype=np.float32)])   # :564-565
        sp = root / self.shortest_paths_path if self.shortest_paths_path else None
        if sp is not None and sp.exists() and (self.hparams['use_position'] or self.hparams['use_neighborhood']):
            self.graph.set_hop_table(load_hop_table(sp, n_nodes))

    def _cache(self):
        if not self.similarities_path:
            return None
        return SimilarityCache(PROJECT_ROOT / self.similarities_path, self.hparams, write=bool(self.hparams.get('b200_write_caches', True)))

    def prepare_data(self, seed=None):
        """SubGNN.py:1024-1063 on the GPU (components, border sets, walks, DTW / SP similarities, anchors); every cached
        product under <task>/similarities/ is read / written with the reference's file names (formats.SimilarityCache)."""
        from . import prepare as prep
        if self.engine is not None:                      # Lightning calls prepare_data once; from_prepared models are ready
            return
        seed = self.hparams.get('seed', 0) if seed is None else seed
        splits = [s for s in ('train', 'val') if len(self.sub_G[s])]
        self.sim_cache = self._cache()
        prepared = prep.prepare(self.hparams, self.graph, self.sub_G, self.sub_G_label, self.embeddings, seed=seed, splits=splits,
                                num_classes=self.num_classes, cache=self.sim_cache, multilabel=self.multilabel)
        eng = Engine(self.hparams, prepared, device=self.device_, graph=self.graph, seed=seed)
        eng.init_parameters(seed)
        self.prepared = prepared
        self._attach(eng)

    def prepare_test_data(self, seed=None):
        """SubGNN.py:994-1022: same products for the test split; P-border and structure anchors are shared with training."""
        from . import prepare as prep
        if 'test' in self.engine.prepared['cc_ids']:
            return
        seed = self.hparams.get('seed', 0) if seed is None else seed
        q = prep.prepare(self.hparams, self.graph, self.sub_G, self.sub_G_label, self.embeddings, seed=seed, splits=['test'],
                         num_classes=self.num_classes, cache=getattr(self, 'sim_cache', None) or self._cache(), multilabel=self.multilabel,
                         shared=self.engine.prepared)
        self.engine.add_split('test', q)

    # ---- forward / steps -----------------------------------------------------------------------------
    def forward(self, dataset_type, *unused, subgraph_idx=None, **kw):
        """SubGNN.py:225.  Only ``dataset_type`` and ``subgraph_idx`` are consumed: component ids, anchors and
        similarities of the split are already resident on the device (the reference's remaining positional
        arguments are accepted and ignored)."""
        if subgraph_idx is None:
            # reference positional order: (.., subgraph_ids, cc_ids, subgraph_idx, NP_sim, I_S_sim, B_S_sim)
            subgraph_idx = unused[8]
        idx = subgraph_idx.reshape(-1).cpu().numpy()
        params = [self.get_parameter(n) for n in self._param_names]
        return _EngineFunction.apply(self, dataset_type, idx, self.training, *params)

    def _loss(self, logits, labels):
        if self.multilabel:
            return nn.functional.binary_cross_entropy_with_logits(logits.squeeze(1), labels.type_as(logits))
        return nn.functional.cross_entropy(logits, labels)

    def training_step(self, train_batch, batch_idx=0):
        """SubGNN.py:317-348 (autograd path: usable with any torch optimizer / trainer loop)."""
        logits = self.forward('train', subgraph_idx=train_batch['subgraph_idx'])
        labels = train_batch['label'].to(logits.device)
        labels = labels.reshape(logits.shape[0], -1) if self.multilabel else labels.reshape(-1)
        loss = self._loss(logits, labels)
        acc = calc_accuracy(logits.detach(), labels, self.multilabel)
        return {'loss': loss, 'log': {'train_loss': loss, 'train_acc': acc}}

    def training_step_fused(self, train_batch, use_graph=True, sync_loss=False):
        """Whole optimisation step on the device (forward, loss, backward, clip_grad_norm_, Adam) — the fast path.
        ``train_batch['subgraph_idx']`` may be a host tensor; it is the only per-step input.  The returned loss is a device
        tensor of an asynchronously enqueued step; with ``sync_loss`` the call waits for the step and returns the loss as a
        host tensor (copied by the D2H node that ends the step graph), so ``float(loss)`` costs no further transfer."""
        idx = train_batch['subgraph_idx']
        if idx.is_cuda:
            idx = idx.cpu()
        loss = self.engine.train_step(idx, use_graph=use_graph)
        if sync_loss:
            torch.cuda.current_stream().synchronize()
            return {'loss': self.engine._last_ctx.loss_host}      # pinned host tensor written by the graph's D2H node (valid until the next step)
        return {'loss': loss}

    def val_test_step(self, batch, batch_idx=0, is_test=False):
        """SubGNN.py:350-391."""
        k = 'test' if is_test else 'val'
        with torch.no_grad():
            idx = batch['subgraph_idx'].reshape(-1).cpu().numpy()
            logits, loss = self.engine.forward(k, idx, training=False)
        labels = batch['label'].to(logits.device)
        labels = labels.reshape(len(idx), -1) if self.multilabel else labels.reshape(-1)
        logits = logits.clone()
        acc = calc_accuracy(logits, labels, self.multilabel)
        macro_f1 = calc_f1(logits, labels, 'macro', self.multilabel)
        return {k + '_loss': loss.clone().squeeze(), k + '_acc': acc, k + '_macro_f1': macro_f1, k + '_logits': logits, k + '_labels': labels}

    def validation_step(self, val_batch, batch_idx=0):
        return self.val_test_step(val_batch, batch_idx, is_test=False)

    def test_step(self, test_batch, batch_idx=0):
        return self.val_test_step(test_batch, batch_idx, is_test=True)

    # ---- epoch end (SubGNN.py:408-504): host-side sklearn metrics over the gathered logits -------------------
    def _epoch_end(self, outputs, k):
        logits = torch.cat([x[k + '_logits'] for x in outputs], dim=0)
        labels = torch.cat([x[k + '_labels'] for x in outputs], dim=0)
        logs = epoch_metrics(logits, labels, self.multilabel, k)
        avg_loss = torch.stack([x[k + '_loss'] for x in outputs]).mean().cpu()
        avg_acc = torch.stack([x[k + '_acc'] for x in outputs]).mean()
        avg_macro_f1 = torch.stack([x[k + '_macro_f1'] for x in outputs]).mean()
        head = {k + '_loss': avg_loss, k + '_micro_f1': logs.pop('micro_f1'), k + '_macro_f1': logs.pop('macro_f1'), k + '_acc': logs.pop('acc'),
                'avg_%s_acc' % k: avg_acc, ('avg_macro_f1' if k == 'val' else 'test_avg_macro_f1'): avg_macro_f1, k + '_auroc': logs.pop('auroc')}
        head.update(logs)
        return avg_loss, head

    def validation_epoch_end(self, outputs):
        avg_loss, logs = self._epoch_end(outputs, 'val')
        if self.hparams.get('resample_anchor_patches', False):                                  # SubGNN.py:449-457
            self.resample_anchor_patches()
        self.metric_scores.append(logs)                                                        # keep track for optuna (:459)
        return {'avg_val_loss': avg_loss, 'log': logs}

    def test_epoch_end(self, outputs):
        avg_loss, logs = self._epoch_end(outputs, 'test')
        self.test_results = logs
        return {'avg_test_loss': avg_loss, 'log': logs}

    def resample_anchor_patches(self):
        """SubGNN.py:449-457: draw fresh N / P / S anchors for train + val (similarities come from the cache / hop table)."""
        from . import prepare as prep
        self._resample_round = getattr(self, '_resample_round', 0) + 1
        old = self.engine.prepared
        splits = [s for s in ('train', 'val') if s in old['cc_ids']]
        sub_G, labels = {s: old['sub_G'][s] for s in splits}, {s: old['labels'][s] for s in splits}
        from .formats import MemoryCache
        q = prep.prepare(self.hparams, self.engine.graph, sub_G, labels, old['embeddings'], seed=self.engine.seed + 7717 * self._resample_round,
                         splits=splits, num_classes=self.num_classes, multilabel=self.multilabel, cache=MemoryCache(self.hparams, old))
        self.engine.rebind(q)

    # ---- loaders (SubGNN.py:1116-1151): batches carry indices and labels only ---------------------------
    def _loader(self, split, shuffle):
        labels = torch.as_tensor(np.asarray(self.engine.prepared['labels'][split]))
        B = self.hparams['batch_size']
        return IndexLoader(labels, B, shuffle, drop_last=shuffle and B <= len(labels))           # drop_last: SubGNN.py:1126

    def train_dataloader(self):
        return self._loader('train', True)

    def val_dataloader(self):
        return self._loader('val', False)

    def test_dataloader(self):
        self.prepare_test_data()                                                               # SubGNN.py:1146
        return self._loader('test', False)

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.hparams['learning_rate'])               # SubGNN.py:1156-1161

    def backward(self, trainer, loss, optimizer, optimizer_idx):
        loss.backward(retain_graph=True)                                                        # SubGNN.py:1163-1164


class IndexLoader:
    """Re-iterable stand-in for DataLoader(SubgraphDataset, collate_fn=_pad_collate) (SubGNN.py:1068-1151): a batch is
    {'subgraph_idx': (B, 1) int64, 'label': (B,) or (B, K)} — component ids, border sets and similarity slabs stay on the device."""

    def __init__(self, labels, batch_size, shuffle, drop_last):
        self.labels, self.batch_size, self.shuffle, self.drop_last = labels, batch_size, shuffle, drop_last

    def __len__(self):
        n, B = len(self.labels), self.batch_size
        return n // B if self.drop_last else (n + B - 1) // B

    def __iter__(self):
        n, B = len(self.labels), self.batch_size
        order = torch.randperm(n) if self.shuffle else torch.arange(n)
        for b in range(len(self)):
            idx = order[b * B:(b + 1) * B]
            yield {'subgraph_idx': idx.view(-1, 1), 'label': self.labels[idx]}


# ---- metrics (subgraph_utils.py:93-124, SubGNN.py:408-504): sklearn on the host, as in the reference ----------------------
def _pred(logits, multilabel):
    return (torch.sigmoid(logits) > 0.5) if multilabel else torch.argmax(logits, dim=-1)


def calc_f1(logits, labels, avg_type='macro', multilabel=False):
    from sklearn.metrics import f1_score
    return torch.tensor([f1_score(labels.cpu().numpy(), _pred(logits, multilabel).cpu().numpy(), average=avg_type)])


def calc_accuracy(logits, labels, multilabel=False):
    from sklearn.metrics import accuracy_score
    return torch.tensor([accuracy_score(labels.cpu().numpy(), _pred(logits, multilabel).cpu().numpy())])


def epoch_metrics(logits, labels, multilabel, prefix):
    """micro / macro F1, accuracy, AUROC (+ per-class AUROC keyed '<prefix>_auroc_class_<c>') exactly as
    validation_epoch_end / test_epoch_end compute them."""
    from sklearn.metrics import roc_auc_score
    logits, labels = logits.detach().float().cpu(), labels.detach().cpu()
    out = {'macro_f1': calc_f1(logits, labels, 'macro', multilabel).squeeze(), 'micro_f1': calc_f1(logits, labels, 'micro', multilabel).squeeze(),
           'acc': calc_accuracy(logits, labels, multilabel).squeeze()}
    if multilabel:
        out['auroc'] = roc_auc_score(labels, torch.sigmoid(logits), multi_class='ovr')
    elif len(torch.unique(labels)) == 2:
        out['auroc'] = roc_auc_score(labels, torch.softmax(logits, dim=1)[:, 1])
    else:
        out['auroc'] = roc_auc_score(labels, torch.softmax(logits, dim=1), multi_class='ovr')
    for c in range(logits.shape[1]):
        if multilabel:
            out['%s_auroc_class_%d' % (prefix, c)] = roc_auc_score(labels[:, c], torch.sigmoid(logits)[:, c])
        else:
            out['%s_auroc_class_%d' % (prefix, c)] = roc_auc_score(torch.nn.functional.one_hot(labels, num_classes=logits.shape[1])[:, c], logits[:, c])
    return out
