"""SG_MPN with the reference's interface (SubGNN/subgraph_mpn.py:21-241) on hand-written CUDA.

    SG_MPN(hparams)                     .linear: Linear(2D, D), .linear_position: Linear(D, 1)   (:29-34)
    forward(networkx_graph, sims, cc_ids, cc_embeds, cc_embed_mask, anchor_patches, anchor_embeds, anchor_mask,
            anchors_sim_index) -> (cc_embed (B, C, D), position_struc_out (B, C, A))              (:133-174)

The edge list / similarity lookup / message / scatter-add / property-aware projection are one kernel
(csrc/mpn.cu); the update projection is the tiled GEMM of csrc/gemm.cu.  Differentiable w.r.t. cc_embeds,
anchor_embeds and the four parameters; ``sims`` is data.  No CPU fallback.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._abi import call, ptr, stream_ptr


class _MPNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cc, x, sims, anchor_ids, sim_index, mask, W, b, wp, bp, use_proj):
        R, D = cc.shape
        A = x.shape[1]
        dev = cc.device
        cat = torch.empty((R, 2 * D), dtype=torch.float32, device=dev)
        pos_lin = torch.empty((R, A), dtype=torch.float32, device=dev)
        s_eff = torch.empty((R, A), dtype=torch.float32, device=dev)
        st = stream_ptr()
        call('subgnn_mpn_fwd', ptr(cc), ptr(x), ptr(sims), sims.shape[1], ptr(anchor_ids), ptr(sim_index), ptr(mask), ptr(wp), ptr(bp),
             ptr(cat), ptr(pos_lin), ptr(s_eff), R, A, D, st)
        if use_proj:
            out = torch.empty((R, D), dtype=torch.float32, device=dev)
            call('subgnn_linear_fwd', ptr(cat), 2 * D, None, ptr(W), 2 * D, ptr(b), ptr(out), D, R, D, 2 * D, 1, st)   # relu(W[x;agg]+b) :239
        else:
            out = cat[:, D:].contiguous()                                                                               # :240-241
        ctx.save_for_backward(x, cat, out, s_eff, W, wp)
        ctx.use_proj = use_proj
        return out, pos_lin

    @staticmethod
    def backward(ctx, d_out, d_pos):
        x, cat, out, s_eff, W, wp = ctx.saved_tensors
        R, A, D = x.shape
        dev = x.device
        st = stream_ptr()
        d_out = d_out.contiguous()
        d_pos = d_pos.contiguous()
        dW = torch.zeros_like(W)
        db = torch.zeros(D, dtype=torch.float32, device=dev)
        if ctx.use_proj:
            dpre = (d_out * (out > 0)).contiguous()
            dcat = torch.empty((R, 2 * D), dtype=torch.float32, device=dev)
            call('subgnn_linear_bwd_input', ptr(dpre), D, ptr(W), 2 * D, ptr(dcat), 2 * D, None, R, D, 2 * D, 0, st)
            call('subgnn_linear_bwd_weight', ptr(dpre), D, ptr(cat), 2 * D, None, ptr(dW), 2 * D, ptr(db), R, D, 2 * D, None, st)
        else:
            dcat = torch.zeros((R, 2 * D), dtype=torch.float32, device=dev)
            dcat[:, D:] = d_out
        dx = torch.empty_like(x)
        dwp = torch.zeros_like(wp)
        dbp = torch.zeros(1, dtype=torch.float32, device=dev)
        call('subgnn_mpn_bwd', ptr(x), ptr(s_eff), ptr(dcat), ptr(d_pos), ptr(wp), ptr(dx), ptr(dwp), ptr(dbp), R, A, D, st)
        return dcat[:, :D].contiguous(), dx, None, None, None, None, dW, db, dwp, dbp, None


class SG_MPN(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self.device = torch.device('cuda')
        D = hparams['node_embed_size']
        self.linear = nn.Linear(2 * D, D).to(self.device)
        self.linear_position = nn.Linear(D, 1).to(self.device)

    def forward(self, networkx_graph, sims, cc_ids, cc_embeds, cc_embed_mask, anchor_patches, anchor_embeds, anchor_mask, anchors_sim_index):
        B, C, D = cc_embeds.shape
        A = anchor_patches.shape[2]
        R = B * C
        dev = cc_embeds.device
        mask = anchor_mask.reshape(R, A, -1)[:, :, 0].to(torch.uint8).contiguous()           # :69  first entry of every patch
        if anchors_sim_index is None:
            ids = anchor_patches.reshape(R, A, -1)[:, :, 0].to(torch.int32).contiguous()     # :92  anchor_ids - 1 column
            sidx = None
        else:
            ids = None
            sidx = torch.as_tensor(anchors_sim_index, dtype=torch.int32, device=dev).contiguous()
        out, pos_lin = _MPNFunction.apply(cc_embeds.reshape(R, D).contiguous().float(), anchor_embeds.reshape(R, A, D).contiguous().float(),
                                          sims.reshape(R, -1).contiguous().float().to(dev), ids, sidx, mask, self.linear.weight, self.linear.bias,
                                          self.linear_position.weight.reshape(-1), self.linear_position.bias, bool(self.hparams['use_mpn_projection']))
        if self.hparams.get('norm_pos_struc_embed', False):                                  # :126-129
            pos = F.normalize(pos_lin, p=2, dim=-1)
        else:
            pos = F.relu(pos_lin)
        return out.view(B, C, -1), pos.view(B, C, -1)
