"""subgnn_b200 — B200-native SubGNN subgraph message-passing hot path.

Host side mirrors the reference's Python modules (same function names / argument meaning):
  subgnn_b200.anchor_patch_samplers, subgnn_b200.gamma, subgnn_b200.subgraph_mpn, subgnn_b200.SubGNN
and calls hand-written sm_100a CUDA through the C ABI in include/subgnn_b200.h (ctypes, no torch
types in the signatures).  There is no CPU fallback.
"""
PAD_VALUE = 0  # config.py:8
