"""`relation` graph-type pieces: windowed edge construction and masked edge
attention on the sm_100a kernels.  Host-side mirror of code/model.py:532-611 and :439-471.

Edge order is canonical (dialogue, source j, target i ascending); the reference's CPython-set order is not
reproducible, so the bit-exact contract is on the sorted list."""
import numpy as np
import torch

from . import ops
from ._lib import MMDFNError, call, ptr, stream

I64 = torch.int64
I32 = torch.int32


def edge_perms(l, window_past, window_future):
    """code/model.py:532-550: (j, i) with max(0,j-wp) <= i <= min(l-1,j+wf) (-1 = unbounded), sorted."""
    out = []
    for j in range(l):
        lo = 0 if window_past == -1 else max(0, j - window_past)
        hi = l if window_future == -1 else min(l, j + window_future + 1)
        out.extend((j, i) for i in range(lo, hi))
    return out


def edge_count(l, window_past, window_future):
    j = np.arange(l)
    lo = np.zeros(l, np.int64) if window_past == -1 else np.maximum(0, j - window_past)
    hi = np.full(l, l, np.int64) if window_future == -1 else np.minimum(l, j + window_future + 1)
    return int((hi - lo).sum())


class EdgeSet:
    """Device-side edge list of one batch (int64, canonical order) + CSR-like row pointers per source node."""

    def __init__(self, qmask, geom, window_past, window_future):
        qmask = ops._f32c(qmask)
        T, B, S = qmask.shape
        dev = qmask.device
        self.geom, self.wp, self.wf = geom, int(window_past), int(window_future)
        self.counts = [edge_count(L, self.wp, self.wf) for L in geom.lengths]
        off = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        self.E = int(off[-1])
        self.edge_off = torch.from_numpy(off).to(dev)
        self.edge_index = torch.empty((2, self.E), dtype=I64, device=dev)
        self.edge_type = torch.empty((self.E,), dtype=I64, device=dev)
        self.row_ptr = torch.empty((geom.N + 1,), dtype=I64, device=dev)
        self.node_dia = torch.empty((max(geom.N, 1),), dtype=I32, device=dev)
        call("mmdfn_edges_build", T, B, S, geom.N, self.wp, self.wf, ptr(geom.dia_off, I32), ptr(self.edge_off, I64),
             self.E, ptr(qmask), ptr(self.edge_index, I64), ptr(self.edge_type, I64), ptr(self.row_ptr, I64),
             ptr(self.node_dia, I32), stream())


class EdgeAttnFn(torch.autograd.Function):
    """edge_norm (E,) = masked, window-renormalised softmax_T(M W_att^T) of MaskedEdgeAttention 'attn1'."""

    @staticmethod
    def forward(ctx, M, W_att, edges):
        M, W_att = ops._f32c(M), ops._f32c(W_att)
        T, B, D = M.shape
        geom = edges.geom
        msl = W_att.shape[0]
        if geom.Lmax > msl:
            raise MMDFNError(f"dialogue of length {geom.Lmax} exceeds max_seq_len={msl}")
        ncol = max(geom.Lmax, 1)
        dev = M.device
        s_all = torch.empty((T * B, ncol), device=dev)
        st = stream()
        call("mmdfn_gemm", 0, 1, T * B, ncol, D, 1.0, ptr(M), D, ptr(W_att), D, 0.0, ptr(s_all), ncol, None, 0, st)
        edge_norm = torch.empty((edges.E,), device=dev)
        stat = torch.empty((max(geom.N, 1), 2), device=dev)
        call("mmdfn_edge_attn_fwd", T, B, geom.Lmax, ncol, edges.wp, edges.wf, ptr(geom.dia_off, I32),
             ptr(edges.row_ptr, I64), ptr(s_all), ptr(edge_norm), ptr(stat), st)
        ctx.save_for_backward(M, W_att, s_all, edge_norm, stat)
        ctx.edges = edges
        return edge_norm

    @staticmethod
    def backward(ctx, g):
        M, W_att, s_all, edge_norm, stat = ctx.saved_tensors
        edges, geom = ctx.edges, ctx.edges.geom
        T, B, D = M.shape
        ncol = s_all.shape[1]
        dev = M.device
        g = ops._f32c(g)
        st = stream()
        ds = torch.empty_like(s_all)
        call("mmdfn_edge_attn_bwd", T, B, geom.Lmax, ncol, edges.wp, edges.wf, ptr(geom.dia_off, I32),
             ptr(edges.row_ptr, I64), ptr(s_all), ptr(edge_norm), ptr(stat), ptr(g), ptr(ds), st)
        dM = dW = None
        if ctx.needs_input_grad[0]:
            dM = torch.empty_like(M)
            call("mmdfn_gemm", 0, 0, T * B, D, ncol, 1.0, ptr(ds), ncol, ptr(W_att), D, 0.0, ptr(dM), D, None, 0, st)
        if ctx.needs_input_grad[1]:
            dW = torch.zeros_like(W_att)
            call("mmdfn_gemm", 1, 0, ncol, D, T * B, 1.0, ptr(ds), ncol, ptr(M), D, 0.0, ptr(dW), D, None, 0, st)
        return dM, dW, None


class ScoresDenseFn(torch.autograd.Function):
    """compact edge scores -> the reference's dense (B, max_seq_len, T) tensor"""

    @staticmethod
    def forward(ctx, edge_norm, edges, msl, T):
        geom = edges.geom
        dense = torch.empty((geom.B, msl, T), device=edge_norm.device)
        call("mmdfn_edge_scores_dense", edges.E, geom.B, msl, T, ptr(edges.edge_index, I64), ptr(edges.node_dia, I32),
             ptr(geom.dia_off, I32), ptr(edge_norm), ptr(dense), 0, stream())
        ctx.edges, ctx.msl, ctx.T = edges, msl, T
        return dense

    @staticmethod
    def backward(ctx, gd):
        edges, geom = ctx.edges, ctx.edges.geom
        gd = ops._f32c(gd)
        g = torch.empty((edges.E,), device=gd.device)
        call("mmdfn_edge_scores_dense", edges.E, geom.B, ctx.msl, ctx.T, ptr(edges.edge_index, I64),
             ptr(edges.node_dia, I32), ptr(geom.dia_off, I32), ptr(g), ptr(gd), 1, stream())
        return g, None, None, None


def _window_of(edge_ind, lengths):
    """recover (window_past, window_future) from the reference-style edge lists; raise if they are not windowed"""
    wp = wf = 0
    for e in edge_ind:
        for j, i in e:
            wp, wf = max(wp, int(j) - int(i)), max(wf, int(i) - int(j))
    for e, L in zip(edge_ind, lengths):
        if len(e) != edge_count(L, wp, wf) and len(e) != edge_count(L, -1, -1):
            raise NotImplementedError("MaskedEdgeAttention kernels support the windowed edge sets of edge_perms only")
    return wp, wf


def masked_edge_attention(module, M, lengths, edge_ind, qmask=None):
    """MaskedEdgeAttention.forward (code/model.py:439-471): dense scores (B, max_seq_len, T)."""
    from .modules import _geom_of
    wp, wf = _window_of(edge_ind, lengths)
    geom = _geom_of(lengths, M.device)
    T, B, _ = M.shape
    q = qmask if qmask is not None else torch.zeros((T, B, 1), device=M.device)
    edges = EdgeSet(q, geom, wp, wf)
    en = EdgeAttnFn.apply(M, module.scalar.weight, edges)
    return ScoresDenseFn.apply(en, edges, module.max_seq_len, T)


def batch_graphify(features, qmask, lengths, window_past, window_future, edge_type_mapping, att_model, no_cuda):
    """code/model.py:568-611.  Returns (node_features (N,D), edge_index (2,E) int64, edge_norm (E,), edge_type (E,)
    int64, edge_index_lengths) with edges in canonical sorted order."""
    from .modules import _geom_of
    geom = _geom_of(lengths, features.device)
    edges = EdgeSet(qmask, geom, window_past, window_future)
    edge_norm = EdgeAttnFn.apply(features, att_model.scalar.weight, edges)
    node_features = torch.cat([features[:lengths[j], j, :] for j in range(features.size(1))], dim=0)
    return node_features, edges.edge_index, edge_norm, edges.edge_type, list(edges.counts)
