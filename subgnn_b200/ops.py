"""Thin torch-tensor wrappers over the C ABI (one function per entry point of include/subgnn_b200.h).
Device memory and streams come from torch; no computation happens here."""
import torch

from . import _abi
from ._abi import call, ptr, stream_ptr

DTW_EXACT, DTW_FASTDTW_R1, DTW_EXACT_THREAD = 0, 1, 2      # include/subgnn_b200.h


def _i32(t):
    return t.to(dtype=torch.int32).contiguous()


def walk_full(g, n_walks, walk_len, beta, seed):
    out = torch.empty((n_walks, walk_len), dtype=torch.int32, device=g.device)
    call('subgnn_walk_full', ptr(g.rowptr), ptr(g.col), g.n_nodes, n_walks, walk_len, float(beta), int(seed), ptr(out), stream_ptr())
    return out


def unique_sorted_patches(patches):
    """(P, Lp) ids with PAD/duplicates -> (sorted unique ids left-packed int32 (P, Lp), lengths int32 (P,))."""
    p = patches.to(torch.int64)
    big = torch.iinfo(torch.int64).max
    s, _ = torch.where(p == 0, torch.full_like(p, big), p).sort(dim=1)
    dup = torch.zeros_like(s, dtype=torch.bool)
    dup[:, 1:] = s[:, 1:] == s[:, :-1]
    s, _ = torch.where(dup, torch.full_like(s, big), s).sort(dim=1)
    lens = (s != big).sum(dim=1).to(torch.int32)
    s = torch.where(s == big, torch.zeros_like(s), s).to(torch.int32)
    return s.contiguous(), lens.contiguous()


def walk_patch(g, patches, n_walks, walk_len, beta, border, seed):
    """patches: (P, Lp) device int tensor (PAD=0, duplicates allowed)."""
    sp, lens = unique_sorted_patches(patches.to(g.device))
    P_, Lp = sp.shape
    out = torch.empty((P_, n_walks, walk_len), dtype=torch.int32, device=g.device)
    call('subgnn_walk_patch', ptr(g.rowptr), ptr(g.col), g.n_nodes, ptr(sp), ptr(lens), Lp, P_, n_walks, walk_len, float(beta),
         1 if border else 0, int(seed), ptr(out), stream_ptr())
    return out


def sample_rows(row_ptr, items, width, n_anchors, pad_rule, seed, step, position_stream):
    n_rows = row_ptr.numel() - 1
    out = torch.empty((n_rows, n_anchors), dtype=torch.int32, device=items.device)
    call('subgnn_sample_rows', ptr(row_ptr), ptr(items), n_rows, int(width), n_anchors, 1 if pad_rule else 0, int(seed), int(step),
         1 if position_stream else 0, ptr(out), stream_ptr())
    return out


def border_khop(g, cc_ptr, cc_nodes, k):
    """-> (out_ptr int64 [n_cc+1], ids int32) ragged ascending border sets."""
    n_cc = cc_ptr.numel() - 1
    words = (g.n_nodes + 31) // 32
    bitmaps = torch.empty((max(n_cc, 1), words), dtype=torch.int32, device=g.device)
    counts = torch.zeros(max(n_cc, 1), dtype=torch.int32, device=g.device)
    call('subgnn_border_khop_bitmap', ptr(g.rowptr), ptr(g.col), g.n_nodes, ptr(cc_ptr), ptr(cc_nodes), n_cc, int(k), ptr(bitmaps),
         ptr(counts), stream_ptr())
    out_ptr = torch.zeros(n_cc + 1, dtype=torch.int64, device=g.device)
    out_ptr[1:] = torch.cumsum(counts[:n_cc].to(torch.int64), dim=0)
    total = int(out_ptr[-1].item())
    out = torch.empty(max(total, 1), dtype=torch.int32, device=g.device)
    call('subgnn_border_khop_expand', ptr(bitmaps), g.n_nodes, n_cc, ptr(out_ptr), ptr(out), stream_ptr())
    return out_ptr, out[:total]


def sp_min_dense(hop, cc_ptr, cc_nodes):
    n_rows = cc_ptr.numel() - 1
    N = hop.shape[1]
    out = torch.empty((n_rows, N), dtype=torch.float32, device=hop.device)
    call('subgnn_sp_min_dense', hop.data_ptr(), N, hop.stride(0), ptr(cc_ptr), ptr(cc_nodes), n_rows, ptr(out), stream_ptr())
    return out


def sp_min_gather(hop, cc_ptr, cc_nodes, anchors, anchor_row=None):
    """anchors: int32 (n_lists, A); anchor_row: int32 (n_rows,) or None (identity)."""
    n_rows = cc_ptr.numel() - 1
    A = anchors.shape[-1]
    out = torch.empty((n_rows, A), dtype=torch.float32, device=hop.device)
    call('subgnn_sp_min_gather', hop.data_ptr(), hop.stride(0), ptr(cc_ptr), ptr(cc_nodes), n_rows, ptr(anchors), ptr(anchor_row), A, ptr(out),
         stream_ptr())
    return out


def degree_seq(g, rows, internal):
    rows = _i32(rows.to(g.device))
    n, stride = rows.shape
    seq = torch.empty((n, stride), dtype=torch.int32, device=g.device)
    ln = torch.empty((n,), dtype=torch.int32, device=g.device)
    call('subgnn_degree_seq', ptr(g.rowptr), ptr(g.col), ptr(rows), n, stride, 1 if internal else 0, ptr(seq), ptr(ln), stream_ptr())
    return seq, ln


# upper length bounds of the row buckets of the exact DTW: singletons and pairs stay on the thread-per-pair mapping (a wavefront
# group would carry one or two rows on four lanes), longer components run on the wavefront with G = bound lanes (R rows per lane
# above 32), rows longer than 256 on the thread mapping again
_DTW_EDGES = (2, 4, 8, 16, 32, 64, 128, 256)


def dtw_bucket_plan(max_len_a):
    """[(lo, hi, mode)]: rows of A with lo < length <= hi run in one launch of ``mode`` sized for length hi."""
    plan, lo = [], -1                                           # empty rows ride in the first bucket (the kernels write 0)
    for hi in [b for b in _DTW_EDGES if b < max_len_a] + [max_len_a]:
        plan.append((lo, hi, DTW_EXACT if 2 < hi <= 256 else DTW_EXACT_THREAD))
        lo = hi
    return plan


def dtw_batch(seqA, lenA, seqB, lenB, mode=DTW_FASTDTW_R1, max_len_a=None, max_len_b=None, bucketed=True):
    """(nA, nB) similarities 1 / (1 + DTW) of every (row of A, row of B) pair; 0 for an empty sequence (SubGNN.py:831).
    DTW_EXACT buckets the rows of A by length (index lists built with torch on the device) and launches one wavefront per
    bucket; ``bucketed=False`` runs a single launch sized by the longest row."""
    nA, sA = seqA.shape
    nB, sB = seqB.shape
    if max_len_a is None:
        max_len_a = max(int(lenA.max().item()) if nA else 1, 1)
    if max_len_b is None:
        max_len_b = max(int(lenB.max().item()) if nB else 1, 1)
    out = torch.empty((nA, nB), dtype=torch.float32, device=seqA.device)
    if mode != DTW_EXACT or not bucketed or nA == 0 or nB == 0:
        call('subgnn_dtw_batch', ptr(seqA), ptr(lenA), nA, sA, ptr(seqB), ptr(lenB), nB, sB, max_len_a, max_len_b, int(mode), ptr(out),
             stream_ptr())
        return out
    ln = lenA.to(torch.int64)
    for lo, hi, m in dtw_bucket_plan(max_len_a):
        rows = torch.nonzero((ln > lo) & (ln <= hi)).reshape(-1).to(torch.int32)
        if rows.numel() == 0:
            continue
        call('subgnn_dtw_batch_rows', ptr(seqA), ptr(lenA), ptr(rows), int(rows.numel()), sA, ptr(seqB), ptr(lenB), nB, sB, hi, max_len_b,
             int(m), ptr(out), stream_ptr())
    return out


def hop_table(g):
    """uint8 (N, N) BFS hop table on the device (0 = self / unreachable); also stored on the graph."""
    N = g.n_nodes
    stride = (N + 15) // 16 * 16
    buf = torch.empty((N, stride), dtype=torch.uint8, device=g.device)
    call('subgnn_hop_table', ptr(g.rowptr), ptr(g.col), N, 0, N, ptr(buf), stride, stream_ptr())
    g.hop = buf[:, :N]
    return g.hop
