"""`relation` graph-type pieces: windowed edge construction and
masked edge attention.  Host-side mirror of code/model.py:532-611 and :439-471."""
import torch

from . import ops  # noqa: F401


def edge_perms(l, window_past, window_future):
    """code/model.py:532-550.  Returns the (j, i) pairs with max(0,j-wp) <= i <= min(l-1,j+wf) (-1 = unbounded),
    lexicographically sorted (the reference's CPython-set order is not reproducible)."""
    out = []
    for j in range(l):
        lo = 0 if window_past == -1 else max(0, j - window_past)
        hi = l if window_future == -1 else min(l, j + window_future + 1)
        out.extend((j, i) for i in range(lo, hi))
    return out


def masked_edge_attention(module, M, lengths, edge_ind):
    raise NotImplementedError("relation-path kernels land after the GDF path (SURVEY 8f rank 3)")


def batch_graphify(features, qmask, lengths, window_past, window_future, edge_type_mapping, att_model, no_cuda):
    raise NotImplementedError("relation-path kernels land after the GDF path (SURVEY 8f rank 3)")
