"""Oracle: the per-step SubGNN path in plain torch fp32 on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is also the timed ``cpu_baseline`` of
bench.py (kind "port"): it performs the reference's tensor-op sequence — including the
host-side index construction of SG_MPN.create_edge_index, the materialised (B,C,A,D) anchor
embeddings and the dense (B,C,N) similarity slab — so that its step time is representative of
the reference's CPU path.  Restates
  subgraph_mpn.py:21-241            SG_MPN
  SubGNN.py:60-88                   LSTM walk encoder
  anchor_patch_samplers.py:333-433  get_anchor_patches / embed_anchor_patch / aggregate_structure_anchor_patch
  SubGNN.py:195-312                 run_mpn_layer / forward
  SubGNN.py:317-348, 1156-1164      training_step, Adam, backward  (+ Lightning's clip_grad_norm_)
  SubGNN.py:609-622                 initialize_cc_embeddings
  SubGNN.py:1068-1114, datasets.py:38-57   batch assembly (_pad_collate / __getitem__)
  subgraph_utils.py:213-237         masked_sum
Module / parameter names equal the reference's state_dict keys so weights are interchangeable.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

PAD = 0


class LSTM(nn.Module):
    """SubGNN.py:60-88"""

    def __init__(self, n_features, h, dropout=0.0, num_layers=1, aggregator='last'):
        super().__init__()
        self.aggregator = aggregator
        self.lstm = nn.LSTM(n_features, h, num_layers=num_layers, batch_first=True, dropout=dropout, bidirectional=True)
        self.linear = nn.Linear(2 * h, n_features)

    def forward(self, x):
        out, _ = self.lstm(x)
        if self.aggregator == 'last':
            agg = out[:, -1, :]                     # :83 (reverse direction has seen one token)
        elif self.aggregator == 'sum':
            agg = out.sum(dim=1)                    # :85
        else:
            raise NotImplementedError
        return self.linear(agg)


class SG_MPN(nn.Module):
    """subgraph_mpn.py:21-241 with torch_geometric's propagate written out
    (x_j = x[src]; aggregate = scatter-add over dst; update on all rows)."""

    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        D = hparams['node_embed_size']
        self.linear = nn.Linear(2 * D, D)
        self.linear_position = nn.Linear(D, 1)

    def forward(self, sims, cc_ids, cc_embeds, cc_embed_mask, anchor_patches, anchor_embeds, anchor_mask, anchors_sim_index):
        B, C, D = cc_embeds.shape
        A = anchor_patches.shape[2]
        x = torch.cat([anchor_embeds.reshape(-1, D), cc_embeds.reshape(-1, D)])          # :36-50
        anchor_ids = anchor_patches.contiguous().view(-1, anchor_patches.shape[-1])      # :151
        n_anchor = anchor_ids.shape[0]
        # :52-71 — built from python ranges on the host every call, as in the reference
        src = torch.tensor(range(n_anchor))
        dst = (torch.tensor(range(B * C)) + n_anchor).repeat_interleave(A)
        keep = anchor_mask.reshape(-1, anchor_mask.shape[-1])[:, 0]
        src, dst = src[keep], dst[keep]
        # :73-103
        sims2 = sims.reshape(B * C, -1)
        cc_idx = dst - n_anchor
        if anchors_sim_index is None:
            sim = sims2[cc_idx, (anchor_ids[src, :] - 1).squeeze(-1)]
        else:
            tiled = list(anchors_sim_index) * int(torch.unique(dst).shape[0])
            sim = sims2[cc_idx, torch.tensor(tiled, dtype=torch.long)]
        sim = sim.unsqueeze(-1)
        # :176-241
        msg = sim * x.index_select(0, src)
        agg = torch.zeros_like(x).index_add_(0, dst, msg)
        if self.hparams['use_mpn_projection']:
            upd = F.relu(self.linear(torch.cat([x, agg], dim=1)))
        else:
            upd = agg
        # :105-131
        pos = torch.zeros(B * C * A, D)
        pos[keep] = msg
        pos = self.linear_position(pos.view(-1, A, D)).squeeze(-1)
        if self.hparams.get('norm_pos_struc_embed', False):
            pos = F.normalize(pos, p=2, dim=-1)
        else:
            pos = F.relu(pos)
        return upd[n_anchor:, :].view(B, C, -1), pos.view(B, C, -1)


def masked_sum(v, mask, dim):
    """subgraph_utils.py:213-237"""
    return v.masked_fill(~mask, 0.0).sum(dim=dim)


class OracleSubGNN(nn.Module):
    """The model part of SubGNN.py (forward / training_step / batch assembly) over an in-memory
    ``prepared`` dict (see subgnn_b200/prepared.py for the field list)."""

    def __init__(self, hparams, prepared):
        super().__init__()
        self.hparams = dict(hparams)
        self.p = prepared
        hp = self.hparams
        D = hp['node_embed_size']
        L = hp['n_layers']
        emb = torch.as_tensor(prepared['embeddings'], dtype=torch.float32)
        self.node_embeddings = nn.Embedding.from_pretrained(emb.clone(), freeze=hp['freeze_node_embeds'], padding_idx=PAD)  # SubGNN.py:568
        hid = D
        mk = lambda: nn.ModuleDict({'internal': SG_MPN(hp), 'border': SG_MPN(hp)})
        self.neighborhood_mpns = nn.ModuleList()
        self.position_mpns = nn.ModuleList()
        self.structure_mpns = nn.ModuleList()
        if hp['use_neighborhood']:
            hid += L * 2 * D
            self.neighborhood_mpns.extend(mk() for _ in range(L))
        if hp['use_position']:
            hid += (hp['n_anchor_patches_pos_in'] + hp['n_anchor_patches_pos_out']) * L
            self.position_mpns.extend(mk() for _ in range(L))
        if hp['use_structure']:
            hid += 2 * hp['n_anchor_patches_structure'] * L
            self.structure_mpns.extend(mk() for _ in range(L))
        if hp.get('batch_norm', False):
            raise NotImplementedError('batch_norm: not in any shipped config; outside the oracle')
        self.hid_dim = hid
        self.num_classes = int(prepared['num_classes'])
        self.multilabel = bool(prepared.get('multilabel', False))
        self.lin = nn.Linear(hid, hp['linear_hidden_dim_1'])
        self.lin2 = nn.Linear(hp['linear_hidden_dim_1'], hp['linear_hidden_dim_2'])
        self.lin3 = nn.Linear(hp['linear_hidden_dim_2'], self.num_classes)
        self.lin_dropout = nn.Dropout(p=hp['lin_dropout'])
        self.lin_dropout2 = nn.Dropout(p=hp['lin_dropout'])
        self.loss = nn.BCEWithLogitsLoss() if self.multilabel else nn.CrossEntropyLoss()
        self.lstm = LSTM(D, D, dropout=hp['lstm_dropout'], num_layers=hp['lstm_n_layers'], aggregator=hp['lstm_aggregator'])
        self._t = lambda a: torch.as_tensor(np.asarray(a))
        if hp['trainable_cc']:
            cc0 = self.initialize_cc_embeddings(self._t(prepared['cc_ids']['train'])).detach()
            for name in ('N_I', 'N_B', 'S_I', 'S_B', 'P_I', 'P_B'):          # SubGNN.py:629-635
                setattr(self, 'train_%s_cc_embed' % name, nn.Parameter(cc0.clone()))

    def snapshot_eval_cc_tables(self):
        """prepare_data-time pooling of the val/test components (SubGNN.py:646-668); call after
        loading weights when trainable_cc is on."""
        self._eval_cc_tables = {s: self.initialize_cc_embeddings(self._t(self.p['cc_ids'][s])).detach()
                                for s in self.p['cc_ids'] if s != 'train'}

    # SubGNN.py:609-622
    def initialize_cc_embeddings(self, cc_ids):
        e = self.node_embeddings(cc_ids)
        if self.hparams['cc_aggregator'] == 'sum':
            return e.sum(dim=2)
        if self.hparams['cc_aggregator'] == 'max':
            return e.max(dim=2)[0]
        raise NotImplementedError

    # anchor_patch_samplers.py:333-433
    def get_anchor_patches(self, split, subgraph_idx, cc_ids, cc_mask, layer, channel, inside):
        p, hp = self.p, self.hparams
        B, C, _ = cc_ids.shape
        if channel == 'neighborhood':
            src = p['anchors_neigh_int'] if inside else p['anchors_neigh_border']
            patches = self._t(src[split][layer])[subgraph_idx].squeeze(1)
            embeds = self.node_embeddings(patches)
            mask = (patches != PAD)
            return patches.unsqueeze(-1), mask.unsqueeze(-1), embeds
        if channel == 'position':
            if inside:
                patches = self._t(p['anchors_pos_int'][split][layer])[subgraph_idx].squeeze(1).unsqueeze(1).repeat(1, C, 1)
            else:
                patches = self._t(p['anchors_pos_ext'][layer]).unsqueeze(0).unsqueeze(0).repeat(B, C, 1)
            patches[~cc_mask] = PAD
            embeds = self.node_embeddings(patches)
            mask = (patches != PAD)
            return patches.unsqueeze(-1), mask.unsqueeze(-1), embeds
        if channel == 'structure':
            patches, _, int_rw, bor_rw = p['anchors_structure'][layer]
            patches = self._t(patches)
            rw = self._t(int_rw if inside else bor_rw)
            A, W, T = rw.shape
            walk_embeds = self.node_embeddings(rw).view(A * W, T, -1)
            embeds = self.lstm(walk_embeds).view(A, W, -1).sum(dim=1)
            patches = patches.unsqueeze(0).unsqueeze(0).repeat(B, C, 1, 1)
            patches[~cc_mask] = PAD
            mask = (patches != PAD)
            embeds = embeds.unsqueeze(0).unsqueeze(0).repeat(B, C, 1, 1)
            embeds[~cc_mask] = PAD
            return patches, mask, embeds
        raise Exception('An invalid channel has been entered.')

    def run_mpn_layer(self, split, mpn, subgraph_idx, cc_ids, cc_embeds, cc_mask, sims, layer, channel, inside):
        patches, mask, embeds = self.get_anchor_patches(split, subgraph_idx, cc_ids, cc_mask, layer, channel, inside)
        sim_index = self.p['anchors_structure'][layer][1] if channel == 'structure' else None
        return mpn(sims, cc_ids, cc_embeds, cc_mask, patches, embeds, mask, sim_index)

    # SubGNN.py:225-312
    def forward(self, split, batch):
        hp = self.hparams
        cc_ids, sub_idx = batch['cc_ids'], batch['subgraph_idx']
        init = self.initialize_cc_embeddings(cc_ids)
        if not hp['trainable_cc']:
            ch = {k: init.clone() for k in ('N_I', 'N_B', 'P_I', 'P_B', 'S_I', 'S_B')}
        elif split != 'train':
            # SubGNN.py:659-668: val/test channel tables are plain tensors pooled once in prepare_data
            # (never refreshed when trainable_cc, :449) -> a snapshot of the prepare-time embeddings
            snap = self._eval_cc_tables[split]
            ch = {k: torch.index_select(snap, 0, sub_idx.squeeze(-1)) for k in ('N_I', 'N_B', 'P_I', 'P_B', 'S_I', 'S_B')}
        else:
            ch = {k: torch.index_select(getattr(self, 'train_%s_cc_embed' % k), 0, sub_idx.squeeze(-1))
                  for k in ('N_I', 'N_B', 'P_I', 'P_B', 'S_I', 'S_B')}
        cc_mask = (cc_ids != PAD)[:, :, 0]
        outs = []
        for l in range(hp['n_layers']):
            if hp['use_neighborhood']:
                ch['N_I'], _ = self.run_mpn_layer(split, self.neighborhood_mpns[l]['internal'], sub_idx, cc_ids, ch['N_I'], cc_mask, batch['NP_sim'], l, 'neighborhood', True)
                ch['N_B'], _ = self.run_mpn_layer(split, self.neighborhood_mpns[l]['border'], sub_idx, cc_ids, ch['N_B'], cc_mask, batch['NP_sim'], l, 'neighborhood', False)
                outs += [ch['N_I'], ch['N_B']]
            if hp['use_position']:
                ch['P_I'], pi = self.run_mpn_layer(split, self.position_mpns[l]['internal'], sub_idx, cc_ids, ch['P_I'], cc_mask, batch['NP_sim'], l, 'position', True)
                ch['P_B'], pb = self.run_mpn_layer(split, self.position_mpns[l]['border'], sub_idx, cc_ids, ch['P_B'], cc_mask, batch['NP_sim'], l, 'position', False)
                outs += [pi, pb]
            if hp['use_structure']:
                ch['S_I'], si = self.run_mpn_layer(split, self.structure_mpns[l]['internal'], sub_idx, cc_ids, ch['S_I'], cc_mask, batch['I_S_sim'], l, 'structure', True)
                ch['S_B'], sb = self.run_mpn_layer(split, self.structure_mpns[l]['border'], sub_idx, cc_ids, ch['S_B'], cc_mask, batch['B_S_sim'], l, 'structure', False)
                outs += [si, sb]
        allcc = torch.cat([init] + outs, dim=-1)
        sub = masked_sum(allcc, cc_mask.unsqueeze(-1), dim=1)
        h = self.lin_dropout(F.relu(self.lin(sub)))
        h = self.lin_dropout2(F.relu(self.lin2(h)))
        return self.lin3(h)

    # datasets.py:38-57 + SubGNN.py:1068-1114
    def make_batch(self, split, indices):
        p = self.p
        idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
        items = []
        for i in idx.tolist():
            cc = self._t(p['cc_ids'][split])[i]
            np_sim = self._t(p['NP_sim'][split])[i] if p.get('NP_sim') is not None else None
            i_s = self._t(p['I_S_sim'][split])[i] if p.get('I_S_sim') is not None else None
            b_s = self._t(p['B_S_sim'][split])[i] if p.get('B_S_sim') is not None else None
            lab = self._t(p['labels'][split])[i]
            items.append((cc, np_sim, i_s, b_s, torch.LongTensor([i]), lab))
        cc, np_sim, i_s, b_s, ids, labs = zip(*items)
        cc = torch.stack(cc)
        B, C, _ = cc.shape
        flat = cc.view(B * C, -1)
        keep = flat.abs().sum(dim=0) != 0                               # :1106-1110 trim all-PAD columns
        cc = flat[:, keep].view(B, C, -1)
        st = lambda t: None if t[0] is None else torch.stack(t)
        return {'cc_ids': cc, 'subgraph_idx': torch.stack(ids), 'label': torch.stack(labs),
                'NP_sim': st(np_sim), 'I_S_sim': st(i_s), 'B_S_sim': st(b_s)}

    def loss_from_logits(self, logits, labels):
        if self.multilabel:
            return self.loss(logits.squeeze(1), labels.type_as(logits))
        return self.loss(logits, labels)

    def training_step(self, batch):
        logits = self.forward('train', batch)
        return self.loss_from_logits(logits, batch['label']), logits


def train_steps(model, optimizer, batches, grad_clip, anomaly=False):
    """Lightning 0.7.1's inner loop for this model: step -> backward(retain_graph) -> clip -> Adam
    (SubGNN.py:1156-1164, train_config.py:109-158 gradient_clip_val)."""
    losses = []
    with torch.autograd.set_detect_anomaly(anomaly):
        for b in batches:
            loss, _ = model.training_step(b)
            optimizer.zero_grad()
            loss.backward(retain_graph=True)
            if grad_clip and grad_clip > 0:
                torch.nn.utils.clip_grad_norm_(model.parameters(), grad_clip)
            optimizer.step()
            losses.append(float(loss.detach()))
    return losses
