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
        self.node_spk, self.S = None, S
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
    _node_speakers(edges, qmask)
    edge_norm = EdgeAttnFn.apply(features, att_model.scalar.weight, edges)
    node_features = torch.cat([features[:lengths[j], j, :] for j in range(features.size(1))], dim=0)
    edges.edge_index._mmdfn_edges = edges          # lets RGCNConv / GraphConv recover the windowed structure
    return node_features, edges.edge_index, edge_norm, edges.edge_type, list(edges.counts)


# ------------------------------------------------------------------------------------------------------------
# a12: GraphNetwork = RGCNConv -> GraphConv (code/model.py:675-715).  The two layers are torch-geometric 1.4.3
# modules in the reference (not vendored / not installed): parameter names, shapes and forward semantics below
# follow PyG 1.4.3 (RGCNConv: basis, att, root, bias; GraphConv: weight, lin.{weight,bias}) -- parity unpinned.
# ------------------------------------------------------------------------------------------------------------
def _gemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta=0.0, bias=None):
    call("mmdfn_gemm", int(ta), int(tb), M, N, K, 1.0, A, lda, B, ldb, float(beta), C, ldc, bias, 0, stream())


def _node_speakers(edges, qmask):
    if getattr(edges, "node_spk", None) is None:
        q = ops._f32c(qmask)
        T, B, S = q.shape
        edges.node_spk = torch.empty((max(edges.geom.N, 1),), dtype=I32, device=q.device)
        edges.S = S
        call("mmdfn_node_speakers", B, S, ptr(edges.geom.dia_off, I32), ptr(q), ptr(edges.node_spk, I32), stream())
    return edges.node_spk


class RGCNConvFn(torch.autograd.Function):
    """out_i = sum_{j->i} norm_e (x_j W_{type_e}) + x_i root + bias,  W_r = sum_b att[r,b] basis[b]."""

    @staticmethod
    def forward(ctx, x, edge_norm, basis, att, root, bias, edges):
        x, edge_norm, basis, att, root, bias = (ops._f32c(t) for t in (x, edge_norm, basis, att, root, bias))
        geom = edges.geom
        N, fin = x.shape
        nb, _, fout = basis.shape
        R = att.shape[0]
        dev = x.device
        basisT = torch.empty((nb, fout, fin), device=dev)
        call("mmdfn_transpose_batched", nb, fin, fout, ptr(basis), ptr(basisT), stream())
        Wt = torch.empty((R, fout * fin), device=dev)                   # Wt[r, o, i] = W_r[i, o]
        _gemm(0, 0, R, fout * fin, nb, ptr(att), nb, ptr(basisT), fout * fin, ptr(Wt), fout * fin)
        xw = torch.empty((N, R * fout), device=dev)                     # xw[n, r, o] = (x W_r)[n, o]
        _gemm(0, 1, N, R * fout, fin, ptr(x), fin, ptr(Wt), fin, ptr(xw), R * fout)
        out = torch.empty((N, fout), device=dev)
        _gemm(0, 0, N, fout, fin, ptr(x), fin, ptr(root), fout, ptr(out), fout, bias=ptr(bias))
        call("mmdfn_rgcn_aggregate_fwd", N, fout, R, edges.S, edges.wp, edges.wf, ptr(geom.dia_off, I32),
             ptr(edges.node_dia, I32), ptr(edges.node_spk, I32), ptr(edges.row_ptr, I64), ptr(xw), ptr(edge_norm), ptr(out),
             stream())
        ctx.save_for_backward(x, edge_norm, basisT, att, root, Wt, xw)
        ctx.edges = edges
        return out

    @staticmethod
    def backward(ctx, dout):
        x, edge_norm, basisT, att, root, Wt, xw = ctx.saved_tensors
        edges, geom = ctx.edges, ctx.edges.geom
        dout = ops._f32c(dout)
        N, fin = x.shape
        nb, fout, _ = basisT.shape
        R = att.shape[0]
        dev = x.device
        dxw = torch.empty_like(xw)
        dnorm = torch.empty_like(edge_norm)
        call("mmdfn_rgcn_aggregate_bwd", N, fout, R, edges.S, edges.wp, edges.wf, ptr(geom.dia_off, I32),
             ptr(edges.node_dia, I32), ptr(edges.node_spk, I32), ptr(edges.row_ptr, I64), ptr(xw), ptr(edge_norm), ptr(dout),
             ptr(dxw), ptr(dnorm), stream())
        d_root = torch.empty_like(root)
        _gemm(1, 0, fin, fout, N, ptr(x), fin, ptr(dout), fout, ptr(d_root), fout)
        d_bias = torch.empty((fout,), device=dev)
        call("mmdfn_colsum", N, fout, ptr(dout), fout, 0.0, ptr(d_bias), stream())
        dx = torch.empty_like(x)
        _gemm(0, 1, N, fin, fout, ptr(dout), fout, ptr(root), fout, ptr(dx), fin)                  # dout root^T
        _gemm(0, 0, N, fin, R * fout, ptr(dxw), R * fout, ptr(Wt), fin, ptr(dx), fin, beta=1.0)    # + dxw Wt
        dWt = torch.empty_like(Wt)
        _gemm(1, 0, R * fout, fin, N, ptr(dxw), R * fout, ptr(x), fin, ptr(dWt), fin)
        d_att = torch.empty_like(att)
        _gemm(0, 1, R, nb, fout * fin, ptr(dWt), fout * fin, ptr(basisT), fout * fin, ptr(d_att), nb)
        d_basisT = torch.empty_like(basisT)
        _gemm(1, 0, nb, fout * fin, R, ptr(att), nb, ptr(dWt), fout * fin, ptr(d_basisT), fout * fin)
        d_basis = torch.empty((nb, fin, fout), device=dev)
        call("mmdfn_transpose_batched", nb, fout, fin, ptr(d_basisT), ptr(d_basis), stream())
        return dx, dnorm, d_basis, d_att, d_root, d_bias, None


class GraphConvFn(torch.autograd.Function):
    """out_i = sum_{j->i} (x W)_j + Linear(x_i)   (aggr='add', no edge weights)."""

    @staticmethod
    def forward(ctx, x, weight, lin_w, lin_b, edges):
        x, weight, lin_w, lin_b = (ops._f32c(t) for t in (x, weight, lin_w, lin_b))
        geom = edges.geom
        N, fin = x.shape
        fout = weight.shape[1]
        dev = x.device
        h = torch.empty((N, fout), device=dev)
        _gemm(0, 0, N, fout, fin, ptr(x), fin, ptr(weight), fout, ptr(h), fout)
        out = torch.empty((N, fout), device=dev)
        _gemm(0, 1, N, fout, fin, ptr(x), fin, ptr(lin_w), fin, ptr(out), fout, bias=ptr(lin_b))
        call("mmdfn_window_sum", N, fout, edges.wf, edges.wp, ptr(geom.dia_off, I32), ptr(edges.node_dia, I32), ptr(h),
             ptr(out), 1, stream())
        ctx.save_for_backward(x, weight, lin_w)
        ctx.edges = edges
        return out

    @staticmethod
    def backward(ctx, dout):
        x, weight, lin_w = ctx.saved_tensors
        edges, geom = ctx.edges, ctx.edges.geom
        dout = ops._f32c(dout)
        N, fin = x.shape
        fout = weight.shape[1]
        dev = x.device
        dh = torch.empty((N, fout), device=dev)
        call("mmdfn_window_sum", N, fout, edges.wp, edges.wf, ptr(geom.dia_off, I32), ptr(edges.node_dia, I32), ptr(dout),
             ptr(dh), 0, stream())
        d_weight = torch.empty_like(weight)
        _gemm(1, 0, fin, fout, N, ptr(x), fin, ptr(dh), fout, ptr(d_weight), fout)
        dx = torch.empty_like(x)
        _gemm(0, 1, N, fin, fout, ptr(dh), fout, ptr(weight), fout, ptr(dx), fin)                 # dh W^T
        _gemm(0, 0, N, fin, fout, ptr(dout), fout, ptr(lin_w), fin, ptr(dx), fin, beta=1.0)       # + dout lin_w
        d_lin_w = torch.empty_like(lin_w)
        _gemm(1, 0, fout, fin, N, ptr(dout), fout, ptr(x), fin, ptr(d_lin_w), fin)
        d_lin_b = torch.empty((fout,), device=dev)
        call("mmdfn_colsum", N, fout, ptr(dout), fout, 0.0, ptr(d_lin_b), stream())
        return dx, d_weight, d_lin_w, d_lin_b, None


def _edges_of(edge_index):
    edges = getattr(edge_index, "_mmdfn_edges", None)
    if edges is None:
        raise NotImplementedError("edge tensors must come from mmdfn_b200.relation.batch_graphify (windowed edge sets)")
    return edges


class RGCNConv(torch.nn.Module):
    """torch_geometric.nn.RGCNConv(in, out, num_relations, num_bases) of PyG 1.4.3 (same parameter names / shapes)."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases, root_weight=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        self.basis = torch.nn.Parameter(torch.Tensor(num_bases, in_channels, out_channels))
        self.att = torch.nn.Parameter(torch.Tensor(num_relations, num_bases))
        self.root = torch.nn.Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.Tensor(out_channels))
        bound = 1.0 / (num_bases * in_channels) ** 0.5            # PyG: uniform(num_bases * in_channels, tensor)
        for p in (self.basis, self.att, self.root, self.bias):
            p.data.uniform_(-bound, bound)

    def forward(self, x, edge_index, edge_type, edge_norm=None, size=None):
        """`edge_type` must be the tensor batch_graphify produced for `edge_index`: the kernels recompute each edge's
        type from the node speakers (2 (S spk_j + spk_i) + [j >= i], code/model.py:599-606) instead of reading it."""
        edges = _edges_of(edge_index)
        if edge_type is not None and edge_type is not edges.edge_type and (
                edge_type.shape != edges.edge_type.shape or not torch.equal(edge_type, edges.edge_type)):
            raise NotImplementedError("RGCNConv: edge_type differs from the speaker/temporal relation types of this edge set")
        if edge_norm is None:
            edge_norm = torch.ones((edges.E,), device=x.device)
        return RGCNConvFn.apply(x, edge_norm, self.basis, self.att, self.root, self.bias, edges)


class GraphConv(torch.nn.Module):
    """torch_geometric.nn.GraphConv(in, out, aggr='add') of PyG 1.4.3."""

    def __init__(self, in_channels, out_channels, aggr='add', bias=True):
        super().__init__()
        if aggr != 'add':
            raise NotImplementedError("only aggr='add'")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.Tensor(in_channels, out_channels))
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=bias)
        bound = 1.0 / in_channels ** 0.5
        self.weight.data.uniform_(-bound, bound)
        # PyG 1.4.3 GraphConv.reset_parameters(): uniform(in_channels, weight), then lin.reset_parameters() -- a SECOND
        # draw for lin (nn.Linear already drew once in its constructor); mirrored so that later modules see the same
        # RNG stream.  (Recalled from the published PyG source; PyG is not vendored: init parity on this path is unpinned.)
        self.lin.reset_parameters()

    def forward(self, x, edge_index, edge_weight=None, size=None):
        if edge_weight is not None:
            raise NotImplementedError("GraphNetwork calls conv2 without edge weights (code/model.py:709)")
        return GraphConvFn.apply(x, self.weight, self.lin.weight, self.lin.bias, _edges_of(edge_index))


def attentive_node_features(emotions, seq_lengths, umask, matchatt_layer, no_cuda=False):
    """code/model.py:614-645 (Eq. 4-6 of DialogueGCN): every node attends over the nodes of its own dialogue through
    MatchingAttention('general2').  emotions: ragged (N, D) node features.  Returns the attentive features of the VALID
    rows, ragged (N, D) -- the reference returns the padded (T, B, D) tensor, whose padded rows its only caller
    (`classify_node_features`, :663) drops again; `umask` is implied by `seq_lengths` and only checked for shape."""
    if getattr(matchatt_layer, "att_type", None) != 'general2':
        raise NotImplementedError("attentive_node_features: only MatchingAttention(att_type='general2') (code/model.py:685)")
    if not emotions.is_cuda:
        raise MMDFNError("attentive_node_features needs CUDA tensors: the B200 path has no CPU fallback")
    lengths = [int(x) for x in seq_lengths]
    if umask is not None and (umask.shape[0] != len(lengths) or umask.shape[1] < max(lengths)):
        raise ValueError("umask does not match seq_lengths")
    g = _nodal_geom(lengths, emotions.device)
    q = ops.LinearFn.apply(emotions, matchatt_layer.transform.weight, matchatt_layer.transform.bias)
    return ops.NodalAttnFn.apply(emotions, q, g)


def _nodal_geom(lengths, device, cache={}):
    key = (tuple(lengths), str(device))
    g = cache.get(key)
    if g is None:
        if len(cache) > 64:
            cache.clear()
        g = cache[key] = ops.NodalGeom(lengths, device)
    return g


def classify_node_features(emotions, seq_lengths, umask, matchatt_layer, linear_layer, dropout_layer, smax_fc_layer, nodal_attn,
                           avec, no_cuda=False, mask=None):
    """code/model.py:647-672: (nodal attention ->) relu(linear) -> dropout -> smax_fc -> log_softmax over the ragged (N, .)
    rows.  `mask` (tests only): (N, hidden) uint8 keep mask injected instead of a drawn one."""
    if nodal_attn:
        emotions = attentive_node_features(emotions, seq_lengths, umask, matchatt_layer, no_cuda)
    hidden = ops.LinearFn.apply(emotions, linear_layer.weight, linear_layer.bias)
    p = float(dropout_layer.p)
    if mask is None and dropout_layer.training and p > 0:
        mask = ops.make_mask(tuple(hidden.shape), p, hidden.device)
    hidden = ops.ReluMaskFn.apply(hidden, mask, 1.0 / (1.0 - p) if mask is not None else 1.0)
    hidden = ops.LinearFn.apply(hidden, smax_fc_layer.weight, smax_fc_layer.bias)
    if avec:
        return hidden
    return ops.LogSoftmaxFn.apply(hidden)


class GraphNetwork(torch.nn.Module):
    """code/model.py:675-715: cat([x, conv2(conv1(x))]); with return_feature=False (the text-only DialogueGCN configuration)
    followed by the nodal-attention classifier head (`classify_node_features`)."""

    def __init__(self, num_features, num_classes, num_relations, max_seq_len, hidden_size=64, dropout=0.5, no_cuda=False,
                 use_GCN=False, return_feature=False):
        super().__init__()
        if use_GCN:
            raise NotImplementedError("only use_GCN=False (the GCNLayer1 side branch is an out-of-scope baseline)")
        self.return_feature, self.no_cuda, self.use_GCN = return_feature, no_cuda, use_GCN
        self.conv1 = RGCNConv(num_features, hidden_size, num_relations, num_bases=30)
        self.conv2 = GraphConv(hidden_size, hidden_size)
        if not self.return_feature:
            from .modules import MatchingAttention
            self.matchatt = MatchingAttention(num_features + hidden_size, num_features + hidden_size, att_type='general2')
            self.linear = torch.nn.Linear(num_features + hidden_size, hidden_size)
            self.dropout = torch.nn.Dropout(dropout)
            self.smax_fc = torch.nn.Linear(hidden_size, num_classes)

    def forward(self, x, edge_index, edge_norm, edge_type, seq_lengths, umask, nodal_attn, avec, mask=None):
        out = self.conv1(x, edge_index, edge_type, edge_norm)
        out = self.conv2(out, edge_index)
        emotions = torch.cat([x, out], dim=-1)
        if self.return_feature:
            return emotions
        return classify_node_features(emotions, seq_lengths, umask, self.matchatt, self.linear, self.dropout, self.smax_fc,
                                      nodal_attn, avec, self.no_cuda, mask=mask)
