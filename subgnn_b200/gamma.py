"""gamma with the reference's entry points (SubGNN/gamma.py) on the CUDA kernels of csrc/gamma.cu.

    get_degree_sequence(graph, nodes, degree_dict=None, internal=True) -> list[int]     gamma.py:21-49
    calc_dist(a, b)                                                                     gamma.py:51-52
    calc_dtw(component_degree, patch_degree) -> float                                   gamma.py:54-59
plus the batched forms the prepare pipeline uses (one launch for all rows / all pairs).  ``graph`` may be the
reference's networkx graph (resolved to a cached device CSR) or a DeviceGraph.
"""
import numpy as np
import torch

from . import ops
from .graph import resolve_graph

STRUCTURE_SIMILARITY_MODE = ops.DTW_FASTDTW_R1      # fastdtw(radius=1), the reference default


def get_degree_sequence(graph, nodes, degree_dict=None, internal=True):
    """``degree_dict`` is accepted for signature compatibility; degrees come from the device CSR."""
    g = resolve_graph(graph)
    row = torch.as_tensor(np.asarray(nodes.cpu() if isinstance(nodes, torch.Tensor) else nodes)).reshape(1, -1)
    if row.numel() == 0:
        return []
    seq, ln = ops.degree_seq(g, row.to(g.device), internal)
    return seq[0, :int(ln[0].item())].cpu().tolist()


def degree_sequences(graph, rows, internal=True):
    """batched: rows (n, L) ids with PAD -> (seq int32 (n, L) ascending zero-padded, len int32 (n,)) on the device."""
    g = resolve_graph(graph)
    return ops.degree_seq(g, torch.as_tensor(rows).to(g.device), internal)


def calc_dist(a, b):
    return ((max(a, b) + 1) / (min(a, b) + 1)) - 1


def calc_dtw(component_degree, patch_degree, mode=None):
    """1 / (1 + fastdtw(component, patch, radius=1, dist=calc_dist)); 0 for an empty sequence (SubGNN.py:831)."""
    if len(component_degree) == 0 or len(patch_degree) == 0:
        return 0.0
    dev = torch.device('cuda')
    a = torch.tensor([list(component_degree)], dtype=torch.int32, device=dev)
    b = torch.tensor([list(patch_degree)], dtype=torch.int32, device=dev)
    la = torch.tensor([a.shape[1]], dtype=torch.int32, device=dev)
    lb = torch.tensor([b.shape[1]], dtype=torch.int32, device=dev)
    out = ops.dtw_batch(a, la, b, lb, STRUCTURE_SIMILARITY_MODE if mode is None else mode, a.shape[1], b.shape[1])
    return float(out[0, 0].item())


def dtw_similarities(seq_a, len_a, seq_b, len_b, mode=None):
    """all pairs (SubGNN.py:811-822) -> fp32 (n_a, n_b) on the device."""
    return ops.dtw_batch(seq_a, len_a, seq_b, len_b, STRUCTURE_SIMILARITY_MODE if mode is None else mode)
