"""CPU oracle for the MM-DFN per-dialogue forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mm-dfn_b200/`` may import this file.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may call it, and only as the checker / the timed CPU
baseline -- never as the product path.

It is a from-scratch *functional* restatement (plain ``torch`` CPU fp32 ops for the
floating-point work so that ``autograd`` yields the reference gradients, ``numpy``
for integer/index work) of the reference algorithm, function by function.  All
``file:line`` cites are relative to ``/root/reference/``.

Parity status: PINNED.  The reference has no tests/golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the *unmodified* reference run in
the build container through an import shim (``tests/golden/make_golden.py``) and
committed under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
file against every one of them.  The ``relation``-path graph convolutions
(torch-geometric 1.4.3 RGCNConv/GraphConv, not vendored, not installed) are
restated from the published PyG 1.4.x semantics and are "parity unpinned".

Two flavours exist for the adjacency / graph-conv pieces:
  * ``faithful=True``  -- the reference's own algorithmic structure (dense (3N)^2
    adjacency, per-dialogue Python loops, the two dense diagonal GEMMs).  This is
    what is timed as the CPU baseline.
  * ``faithful=False`` -- the closed form, block by block, never materialising the
    dense matrix; used for parity checks at sizes where the dense form is too slow.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
PI = float(np.pi)
COS_SCALE = 0.99999  # code/model_mm.py:149,165


# --------------------------------------------------------------------------------------
# a1  input projections                                                code/model.py:853-865
# --------------------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """``nn.Linear``: y = x W^T + b  (code/model.py:1065,1094,1129)."""
    y = x.matmul(w.t())
    return y if b is None else y + b


# --------------------------------------------------------------------------------------
# a2  2-layer bidirectional GRU ("lstm_l", "rnn_parties")               code/model.py:866-868
# --------------------------------------------------------------------------------------
def gru_direction(xg: Tensor, w_hh: Tensor, b_hh: Tensor, reverse: bool) -> Tensor:
    """One direction of one ``nn.GRU`` layer given pre-computed input gates.

    xg: (T, B, 3H) = x W_ih^T + b_ih, gate order r, z, n.  h0 = 0, no packing: the
    recurrence runs over every one of the T rows, padding included.
    """
    T, B, H3 = xg.shape
    H = H3 // 3
    h = xg.new_zeros(B, H)
    out: List[Optional[Tensor]] = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        gh = h.matmul(w_hh.t()) + b_hh
        r = torch.sigmoid(xg[t, :, :H] + gh[:, :H])
        z = torch.sigmoid(xg[t, :, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(xg[t, :, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[t] = h
    return torch.stack(out, 0)


def bigru2(x: Tensor, P: Dict[str, Tensor], prefix: str, inter_mask: Optional[Tensor] = None) -> Tensor:
    """``nn.GRU(200, 100, num_layers=2, bidirectional=True)`` on a (T, B, 200) input.

    ``inter_mask`` (T, B, 200), already scaled by 1/(1-p), is the inter-layer dropout of
    train mode (``dropout=p`` ctor arg, code/model.py:866); ``None`` = identity.
    """
    for layer in (0, 1):
        outs = []
        for sfx, rev in (("", False), ("_reverse", True)):
            xg = linear(x, P[f"{prefix}.weight_ih_l{layer}{sfx}"], P[f"{prefix}.bias_ih_l{layer}{sfx}"])
            outs.append(gru_direction(xg, P[f"{prefix}.weight_hh_l{layer}{sfx}"],
                                      P[f"{prefix}.bias_hh_l{layer}{sfx}"], rev))
        x = torch.cat(outs, -1)
        if layer == 0 and inter_mask is not None:
            x = x * inter_mask
    return x


# --------------------------------------------------------------------------------------
# a3  speaker-party ("crn speaker") block                      code/model.py:1070-1090 etc.
# --------------------------------------------------------------------------------------
def speaker_partition(qmask: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Integer part of the crn-speaker block (bit-exact contract).

    qmask: (T, B, S).  For every dialogue b and speaker p, ``idx = nonzero(qmask[:, b, p])``
    ascending (code/model.py:1075).  Returns
      pos (T, B, S) int32 : rank of t inside idx_{b,p}, or -1 where qmask[t,b,p] == 0
      cnt (B, S)    int32 : len(idx_{b,p})
    """
    T, B, S = qmask.shape
    pos = np.full((T, B, S), -1, dtype=np.int32)
    cnt = np.zeros((B, S), dtype=np.int32)
    for b in range(B):
        for p in range(S):
            idx = np.nonzero(qmask[:, b, p])[0]
            pos[idx, b, p] = np.arange(len(idx), dtype=np.int32)
            cnt[b, p] = len(idx)
    return pos, cnt


def party_encode(U: Tensor, qmask: Tensor, P: Dict[str, Tensor], inter_mask=None) -> Tensor:
    """U_p of code/model.py:1070-1088: gather each speaker's utterances to the front of a
    zero-padded length-T sequence, run the shared ``rnn_parties`` BiGRU over the FULL T,
    scatter the first len(idx) outputs back.  U: (T, B, 200), qmask: (T, B, S)."""
    T, B, H = U.shape
    S = qmask.shape[2]
    U_b = U.transpose(0, 1)                       # (B, T, 200)
    q_b = qmask.transpose(0, 1)                   # (B, T, S)
    U_p = U.new_zeros(B, T, H)
    parties = [U.new_zeros(B, T, H) for _ in range(S)]
    idxs = [[torch.nonzero(q_b[b][:, p]).squeeze(-1) for p in range(S)] for b in range(B)]
    for b in range(B):
        for p in range(S):
            ii = idxs[b][p]
            if ii.numel() > 0:
                parties[p][b, :ii.numel()] = U_b[b][ii]
    enc = [bigru2(parties[p].transpose(0, 1), P, "rnn_parties",
                  None if inter_mask is None else inter_mask[p]).transpose(0, 1) for p in range(S)]
    for b in range(B):
        for p in range(S):
            ii = idxs[b][p]
            if ii.numel() > 0:
                U_p[b][ii] = enc[p][b][:ii.numel()]
    return U_p.transpose(0, 1)


# --------------------------------------------------------------------------------------
# a4  ragged pack                                                     code/model.py:553-565
# --------------------------------------------------------------------------------------
def ragged_pack(features: Tensor, lengths: Sequence[int]) -> Tensor:
    """``simple_batch_graphify``: (T, B, D) -> (N, D), dialogue-major."""
    return torch.cat([features[:lengths[j], j, :] for j in range(features.size(1))], 0)


# --------------------------------------------------------------------------------------
# a6  multimodal adjacency                                          code/model_mm.py:122-180
# --------------------------------------------------------------------------------------
def _angular(c: Tensor) -> Tensor:
    return 1.0 - torch.acos(c * COS_SCALE) / PI


def big_adj_dense(feats: Sequence[Tensor], dia_len: Sequence[int], modal_weight: float = 1.0) -> Tensor:
    """Faithful restatement of ``MM_GCN.create_big_adj`` (code/model_mm.py:122-180): dense
    (M*N, M*N) matrix, per-dialogue loop, Gram matrix as a sum of H outer products (:148),
    cross-modal diagonals (:161-172), ``D.mm(adj).mm(D)`` (:176-178)."""
    M = len(feats)
    N = feats[0].shape[0]
    adj = feats[0].new_zeros(M * N, M * N)
    start = 0
    for L in dia_len:
        ar = torch.arange(L)
        for m in range(M):
            t = feats[m][start:start + L]
            nt = t.permute(1, 0) / torch.sqrt(torch.sum(t * t, dim=1))            # (H, L)
            cos = torch.sum(torch.matmul(nt.unsqueeze(2), nt.unsqueeze(1)), dim=0)  # (L, L)
            adj[start + N * m:start + N * m + L, start + N * m:start + N * m + L] = _angular(cos)
        for m in range(M):
            for n in range(M):
                if m == n:
                    continue
                x1 = feats[m][start:start + L]
                x2 = feats[n][start:start + L]
                n1 = x1.permute(1, 0) / torch.sqrt(torch.sum(x1 * x1, dim=1))
                n2 = x2.permute(1, 0) / torch.sqrt(torch.sum(x2 * x2, dim=1))
                cos = torch.sum((n1 * n2).permute(1, 0), dim=1)
                adj[ar + start + N * m, ar + start + N * n] = _angular(cos) * modal_weight
        start += L
    d = adj.sum(1)
    D = torch.diag(torch.pow(d, -0.5))
    return D.mm(adj).mm(D)


def adj_blocks(feats: Sequence[Tensor], dia_len: Sequence[int], modal_weight: float = 1.0):
    """Closed form of the same adjacency in block-compact layout.

    Returns (blocks, diags):
      blocks[i][m]     : (L_i, L_i) normalised in-modal block of dialogue i, modality m
      diags[i][(m, n)] : (L_i,) normalised cross-modal diagonal, m < n (symmetric)
    Row degree d = sum of the in-modal row + the (M-1) cross-modal entries (each row has
    L+M-1 non-zeros).
    """
    M = len(feats)
    blocks, diags = [], []
    start = 0
    for L in dia_len:
        xh = []
        for m in range(M):
            t = feats[m][start:start + L]
            xh.append(t / torch.sqrt(torch.sum(t * t, dim=1, keepdim=True)))
        S_in = [_angular(xh[m].matmul(xh[m].t())) for m in range(M)]
        S_x = {}
        for m in range(M):
            for n in range(m + 1, M):
                S_x[(m, n)] = _angular(torch.sum(xh[m] * xh[n], dim=1)) * modal_weight
        dinv = []
        for m in range(M):
            d = S_in[m].sum(1)
            for n in range(M):
                if n != m:
                    d = d + S_x[(min(m, n), max(m, n))]
            dinv.append(torch.pow(d, -0.5))
        blocks.append([dinv[m][:, None] * S_in[m] * dinv[m][None, :] for m in range(M)])
        diags.append({k: dinv[k[0]] * v * dinv[k[1]] for k, v in S_x.items()})
        start += L
    return blocks, diags


def blocks_to_dense(blocks, diags, dia_len: Sequence[int], M: int = 3) -> Tensor:
    N = int(sum(dia_len))
    adj = blocks[0][0].new_zeros(M * N, M * N)
    start = 0
    for i, L in enumerate(dia_len):
        ar = torch.arange(L)
        for m in range(M):
            adj[start + N * m:start + N * m + L, start + N * m:start + N * m + L] = blocks[i][m]
        for (m, n), v in diags[i].items():
            adj[ar + start + N * m, ar + start + N * n] = v
            adj[ar + start + N * n, ar + start + N * m] = v
        start += L
    return adj


def adj_matmul_blocks(blocks, diags, dia_len: Sequence[int], z: Tensor, M: int = 3) -> Tensor:
    """hi = A_hat z for z (M*N, G) with A_hat in block-compact form."""
    N = int(sum(dia_len))
    out = []
    for m in range(M):
        rows = []
        start = 0
        for i, L in enumerate(dia_len):
            acc = blocks[i][m].matmul(z[N * m + start:N * m + start + L])
            for n in range(M):
                if n != m:
                    acc = acc + diags[i][(min(m, n), max(m, n))][:, None] * z[N * n + start:N * n + start + L]
            rows.append(acc)
            start += L
        out.append(torch.cat(rows, 0))
    return torch.cat(out, 0)


# --------------------------------------------------------------------------------------
# a8  GraphConvolution (GCNII layer, variant=True)                  code/model_GCN.py:176-189
# --------------------------------------------------------------------------------------
def graph_conv(x: Tensor, adj, h0: Tensor, w: Tensor, lamda: float, alpha: float, l: int,
               variant: bool = True, residual: bool = False) -> Tensor:
    """``adj`` is either a dense tensor or a callable z -> A_hat z."""
    theta = math.log(lamda / l + 1)
    hi = adj(x) if callable(adj) else adj.matmul(x)
    if variant:
        support = torch.cat([hi, h0], 1)
        r = (1 - alpha) * hi + alpha * h0
    else:
        support = (1 - alpha) * hi + alpha * h0
        r = support
    out = theta * support.matmul(w) + (1 - theta) * r
    if residual:
        out = out + x
    return out


# --------------------------------------------------------------------------------------
# a7  GCNII_lyc: fcs[0] + K x [LSTM gate step -> GraphConvolution -> ReLU -> dropout -> +q]
#                                                                   code/model_GCN.py:444-488
# --------------------------------------------------------------------------------------
def lstm_cell(x: Tensor, h: Tensor, c: Tensor, w_ih, w_hh, b_ih, b_hh):
    """One ``nn.LSTM`` step, gate order i, f, g, o (code/model_GCN.py:432-434,466)."""
    g = x.matmul(w_ih.t()) + b_ih + h.matmul(w_hh.t()) + b_hh
    G = h.shape[1]
    i, f, gg, o = g[:, :G], g[:, G:2 * G], g[:, 2 * G:3 * G], g[:, 3 * G:]
    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h = torch.sigmoid(o) * torch.tanh(c)
    return h, c


def gcnii_stack(x: Tensor, adj, P: Dict[str, Tensor], prefix: str, nlayers: int, lamda: float, alpha: float,
                reason_flag: bool = True, use_residue: bool = True, masks: Optional[dict] = None,
                collect: Optional[dict] = None) -> Tensor:
    """``GCNII_lyc.forward`` with ``return_feature=True``.

    masks: optional dict of pre-scaled dropout masks {'x': (3N,200), 'h0': (3N,100),
    'layer': [K x (3N,100)]}; None = dropout is the identity.
    """
    mk = masks or {}
    if "x" in mk:
        x = x * mk["x"]                                                   # :453
    h0 = torch.relu(linear(x, P[f"{prefix}.fcs.0.weight"], P[f"{prefix}.fcs.0.bias"]))   # :454
    z = h0 * mk["h0"] if "h0" in mk else h0                               # :456
    h = torch.zeros_like(z)
    c = torch.zeros_like(z)
    for i in range(nlayers):
        if reason_flag:
            q = z
            h, c = lstm_cell(q, h, c, P[f"{prefix}.rnn.weight_ih_l0"], P[f"{prefix}.rnn.weight_hh_l0"],
                             P[f"{prefix}.rnn.bias_ih_l0"], P[f"{prefix}.rnn.bias_hh_l0"])   # :463-467
            z = h
        z = torch.relu(graph_conv(z, adj, h0, P[f"{prefix}.convs.{i}.weight"], lamda, alpha, i + 1))  # :469
        if "layer" in mk:
            z = z * mk["layer"][i]                                        # :470
        if reason_flag:
            z = z + q                                                     # :472
        if collect is not None:
            collect[f"layer{i}"] = z
    if use_residue:
        z = torch.cat([x, z], -1)                                         # :482-483
    return z


# --------------------------------------------------------------------------------------
# f4 (☆)  GCNII: the per-modality deep GCN of graph_type='DeepGCN'   code/model_GCN.py:224-306
# --------------------------------------------------------------------------------------
def gcnii(x: Tensor, dia_len: Sequence[int], P: Dict[str, Tensor], prefix: str, nlayers: int, lamda: float, alpha: float,
          reason_flag: bool = False, masks: Optional[dict] = None) -> Tensor:
    """``GCNII.forward`` with return_feature=True, use_residue=True, new_graph=False: the uni-modal angular-similarity
    adjacency of ``GCNII.create_big_adj`` (:274-297: the in-modal block of the multimodal one, degree = row sum), then
    the same layer loop as GCNII_lyc WITHOUT the in-loop dropout (:263-271) and ONE dropout after it (:273).
    masks: {'x': (N,200), 'h0': (N,100), 'out': (N,100)} pre-scaled; None = identity."""
    mk = dict(masks or {})
    out_mask = mk.pop("out", None)
    mk.pop("layer", None)
    blocks, diags = adj_blocks([x], dia_len, 1.0)
    adj = lambda z: adj_matmul_blocks(blocks, diags, dia_len, z, M=1)
    F_ = gcnii_stack(x, adj, P, prefix, nlayers, lamda, alpha, reason_flag, True, mk)
    if out_mask is not None:
        F_ = torch.cat([F_[:, :200], F_[:, 200:] * out_mask], -1)
    return F_


# --------------------------------------------------------------------------------------
# a5  MM_GCN.forward                                                 code/model_mm.py:77-120
# --------------------------------------------------------------------------------------
def mm_gcn(a: Tensor, v: Tensor, l: Tensor, dia_len: Sequence[int], P: Dict[str, Tensor], prefix: str,
           nlayers: int, lamda: float, alpha: float, reason_flag: bool = True, modal_weight: float = 1.0,
           masks=None, faithful: bool = False, collect: Optional[dict] = None) -> Tensor:
    """(N,200) x3 -> (N,900) = [a_x200 a_g100 | v.. | l..] (use_speaker/use_modal off)."""
    N = a.shape[0]
    if faithful:
        adj = big_adj_dense([a, v, l], dia_len, modal_weight)
    else:
        blocks, diags = adj_blocks([a, v, l], dia_len, modal_weight)
        adj = lambda z: adj_matmul_blocks(blocks, diags, dia_len, z)
        if collect is not None:
            collect["adj_blocks"] = blocks
            collect["adj_diags"] = diags
    X = torch.cat([a, v, l], 0)                                           # :98
    F_ = gcnii_stack(X, adj, P, f"{prefix}.graph_net", nlayers, lamda, alpha, reason_flag, True, masks, collect)
    return torch.cat([F_[:N], F_[N:2 * N], F_[2 * N:3 * N]], -1)          # :113


# --------------------------------------------------------------------------------------
# a9  head + FocalLoss                           code/model.py:1328-1337 ; code/loss.py:14-34
# --------------------------------------------------------------------------------------
def head(feat: Tensor, w: Tensor, b: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    if mask is not None:
        feat = feat * mask
    return torch.log_softmax(linear(torch.relu(feat), w, b), 1)


def focal_loss(log_prob: Tensor, target: Tensor, gamma: float = 0.0, alpha: Optional[Tensor] = None,
               size_average: bool = True) -> Tensor:
    logpt = log_prob.gather(1, target.view(-1, 1)).view(-1)
    pt = logpt.detach().exp()
    if alpha is not None:
        logpt = logpt * alpha.gather(0, target.view(-1))
    loss = -1 * (1 - pt) ** gamma * logpt
    return loss.mean() if size_average else loss.sum()


# --------------------------------------------------------------------------------------
# whole hot path: DialogueGNNModel.forward, base_model='LSTM', graph_type='GDF',
# multi_modal, modals='avl', att_type='concat_subsequently', use_crn_speaker
#                                           code/model.py:1062-1154,1182-1209,1294-1337
# --------------------------------------------------------------------------------------
def forward_gdf(P: Dict[str, Tensor], textf: Tensor, qmask: Tensor, lengths: Sequence[int], acouf: Tensor,
                visuf: Tensor, *, nlayers: int, speaker_weights=(1.0, 1.0, 1.0), lamda: float = 0.5,
                alpha: float = 0.2, reason_flag: bool = True, use_crn_speaker: bool = True,
                modal_weight: float = 1.0, masks: Optional[dict] = None, faithful: bool = False,
                collect: Optional[dict] = None, att_type: str = "concat_subsequently") -> Tensor:
    """Returns log_prob (N, C).  ``masks`` (all optional, pre-scaled by 1/(1-p)):
    'gru_l' (T,B,200), 'gru_p' {'a'|'v'|'l': [S x (T,B,200)]}, 'gcn' (see gcnii_stack),
    'head' (N,900)."""
    mk = masks or {}
    wa, wv, wl = speaker_weights
    U_a = linear(acouf, P["linear_a.weight"], P["linear_a.bias"])
    U_v = linear(visuf, P["linear_v.weight"], P["linear_v.bias"])
    U_l = linear(textf, P["linear_l.weight"], P["linear_l.bias"])
    E_l = bigru2(U_l, P, "lstm_l", mk.get("gru_l"))
    em_a, em_v, em_l = U_a, U_v, E_l
    if use_crn_speaker:
        gp = mk.get("gru_p", {})
        em_a = U_a + wa * party_encode(U_a, qmask, P, gp.get("a"))
        em_v = U_v + wv * party_encode(U_v, qmask, P, gp.get("v"))
        em_l = E_l + wl * party_encode(U_l, qmask, P, gp.get("l"))
    fa, fv, fl = (ragged_pack(e, lengths) for e in (em_a, em_v, em_l))
    if collect is not None:
        collect.update(features_a=fa, features_v=fv, features_l=fl)
    feat = mm_gcn(fa, fv, fl, lengths, P, "graph_model", nlayers, lamda, alpha, reason_flag, modal_weight,
                  mk.get("gcn"), faithful, collect)
    if collect is not None:
        collect["emotions_feat"] = feat
    if att_type == "mfn":
        return mfn_head(feat, lengths, P, mk)
    return head(feat, P["smax_fc.weight"], P["smax_fc.bias"], mk.get("head"))


# --------------------------------------------------------------------------------------
# a10  windowed edge construction ("relation" graph)          code/model.py:532-550,568-611
# --------------------------------------------------------------------------------------
def edge_list(L: int, window_past: int, window_future: int) -> np.ndarray:
    """Sorted (j, i) edge list of one dialogue: max(0,j-wp) <= i <= min(L-1, j+wf); -1 = unbounded.
    The reference emits these in CPython-set order (code/model.py:532-550); the bit-exact
    contract is on the lexicographically sorted list."""
    out = []
    for j in range(L):
        lo = 0 if window_past == -1 else max(0, j - window_past)
        hi = L if window_future == -1 else min(L, j + window_future + 1)
        for i in range(lo, hi):
            out.append((j, i))
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


def speakers_of(qmask: np.ndarray) -> np.ndarray:
    """(T,B,S) -> (T,B) int: first index with qmask == 1 (code/model.py:591-592); -1 if none."""
    eq = (qmask == 1)
    first = eq.argmax(-1)
    return np.where(eq.any(-1), first, -1).astype(np.int64)


def build_edges(qmask: np.ndarray, lengths: Sequence[int], window_past: int, window_future: int):
    """edge_index (2,E) int64 [row0 = j + offset, row1 = i + offset], edge_type (E,) int64
    = 2*(S*spk_j + spk_i) + (j >= i)  (mapping built at code/model.py:974-980), per-dialogue
    edge counts.  Edges sorted by (dialogue, j, i)."""
    S = qmask.shape[2]
    spk = speakers_of(qmask)
    ei, et, counts = [], [], []
    off = 0
    for b, L in enumerate(lengths):
        e = edge_list(L, window_past, window_future)
        counts.append(len(e))
        if len(e):
            ei.append(e + off)
            sj, si = spk[e[:, 0], b], spk[e[:, 1], b]
            et.append(2 * (S * sj + si) + (e[:, 0] >= e[:, 1]).astype(np.int64))
        off += L
    if ei:
        return np.concatenate(ei, 0).T.copy(), np.concatenate(et, 0), counts
    return np.zeros((2, 0), np.int64), np.zeros((0,), np.int64), counts


# --------------------------------------------------------------------------------------
# a11  MaskedEdgeAttention ('attn1')                                  code/model.py:439-471
# --------------------------------------------------------------------------------------
def masked_edge_attention(M_: Tensor, w_att: Tensor, lengths: Sequence[int], window_past: int,
                          window_future: int) -> Tensor:
    """scores (B, max_seq_len, T): softmax over T of Linear(200->max_seq_len, no bias)(M),
    masked (1 on edges, 1e-10 elsewhere), row-renormalised, zeroed off the edges."""
    T, B, _ = M_.shape
    scale = M_.matmul(w_att.t())                                  # (T, B, max_seq_len)
    a = torch.softmax(scale, dim=0).permute(1, 2, 0)              # (B, max_seq_len, T)
    mask = torch.full(a.shape, 1e-10)
    mask01 = torch.zeros(a.shape)
    for b, L in enumerate(lengths):
        e = edge_list(L, window_past, window_future)
        if len(e):
            mask[b, e[:, 0], e[:, 1]] = 1.0
            mask01[b, e[:, 0], e[:, 1]] = 1.0
    ma = a * mask
    return ma / ma.sum(-1, keepdim=True) * mask01


def edge_norms(scores: Tensor, lengths: Sequence[int], window_past: int, window_future: int) -> Tensor:
    """edge_norm (E,) in the same sorted order as ``build_edges`` (code/model.py:590)."""
    out = []
    for b, L in enumerate(lengths):
        e = edge_list(L, window_past, window_future)
        if len(e):
            out.append(scores[b, e[:, 0], e[:, 1]])
    return torch.cat(out, 0) if out else scores.new_zeros(0)


# --------------------------------------------------------------------------------------
# a12  GraphNetwork (RGCNConv -> GraphConv), PyG 1.4.3 semantics -- PARITY UNPINNED
#      call sites: code/model.py:682-683 (ctor), :708-710 (forward)
# --------------------------------------------------------------------------------------
def rgcn_conv(x: Tensor, edge_index: np.ndarray, edge_type: np.ndarray, edge_norm: Tensor, basis: Tensor,
              att: Tensor, root: Tensor, bias: Tensor) -> Tensor:
    """torch-geometric 1.4.3 ``RGCNConv(in, out, R, num_bases)``: W_r = sum_b att[r,b] basis[b];
    message j->i = (x_j W_{type}) * edge_norm; aggr 'add' at target i = edge_index[1];
    out_i = sum + x_i root + bias."""
    nb, fin, fout = basis.shape
    W = att.matmul(basis.reshape(nb, -1)).reshape(-1, fin, fout)
    src = torch.as_tensor(edge_index[0]); dst = torch.as_tensor(edge_index[1])
    et = torch.as_tensor(edge_type)
    msg = torch.bmm(x[src].unsqueeze(1), W[et]).squeeze(1) * edge_norm[:, None]
    out = x.new_zeros(x.shape[0], fout).index_add(0, dst, msg)
    return out + x.matmul(root) + bias


def pyg_graph_conv(x: Tensor, edge_index: np.ndarray, weight: Tensor, lin_w: Tensor, lin_b: Tensor) -> Tensor:
    """torch-geometric 1.4.3 ``GraphConv(in, out)`` (aggr='add'): out_i = sum_{j->i} (x W)_j + Linear(x_i)."""
    src = torch.as_tensor(edge_index[0]); dst = torch.as_tensor(edge_index[1])
    h = x.matmul(weight)
    out = x.new_zeros(x.shape[0], weight.shape[1]).index_add(0, dst, h[src])
    return out + linear(x, lin_w, lin_b)


# --------------------------------------------------------------------------------------
# a13  MMGatedAttention ('general')                                   code/model.py:757-781
# --------------------------------------------------------------------------------------
def mm_gated_attention(a: Tensor, v: Tensor, l: Tensor, P: Dict[str, Tensor], prefix: str = "gatedatt") -> Tensor:
    """Dropout-free restatement: h_m = tanh(W_m x_m); z_mn = sigmoid(w_mn [x_m, x_n, x_m*x_n]);
    out = [h_av | h_al | h_vl]."""
    lin = lambda name, x: linear(x, P[f"{prefix}.{name}.weight"], P[f"{prefix}.{name}.bias"])
    ha, hv, hl = torch.tanh(lin("transform_a", a)), torch.tanh(lin("transform_v", v)), torch.tanh(lin("transform_l", l))
    z_av = torch.sigmoid(lin("transform_av", torch.cat([a, v, a * v], -1)))
    z_al = torch.sigmoid(lin("transform_al", torch.cat([a, l, a * l], -1)))
    z_vl = torch.sigmoid(lin("transform_vl", torch.cat([v, l, v * l], -1)))
    return torch.cat([z_av * ha + (1 - z_av) * hv, z_al * ha + (1 - z_al) * hl, z_vl * hv + (1 - z_vl) * hl], -1)


# --------------------------------------------------------------------------------------
# a13  MFN (att_type='mfn' ablation)                                  code/model_fusion.py:10-120
# --------------------------------------------------------------------------------------
def mfn_lstm_cell(x: Tensor, h: Tensor, c: Tensor, P: Dict[str, Tensor], prefix: str) -> Tuple[Tensor, Tensor]:
    """nn.LSTMCell: gates i, f, g, o = split(W_ih x + b_ih + W_hh h + b_hh); c' = sig(f) c + sig(i) tanh(g); h' = sig(o) tanh(c')."""
    g = linear(x, P[f"{prefix}.weight_ih"], P[f"{prefix}.bias_ih"]) + linear(h, P[f"{prefix}.weight_hh"], P[f"{prefix}.bias_hh"])
    i, f, gg, o = g.chunk(4, dim=-1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    return torch.sigmoid(o) * torch.tanh(c2), c2


def mfn_forward(x: Tensor, P: Dict[str, Tensor], prefix: str = "", masks: Optional[Sequence[Tensor]] = None) -> Tensor:
    """Restatement of MFN.forward (code/model_fusion.py:62-120).  x (T, n, 900) = [l | a | v] (300 each)
    -> (T, n, 400) = [h_l | h_a | h_v | mem].  Three LSTMCells run independently of the memory; per step the window
    cStar = [c_{t-1} | c_t] (600) is re-weighted by a softmax attention (att1), squashed to the memory proposal cHat
    (att2, tanh), and two sigmoid gates computed from [attended | mem_{t-1}] (gamma1, gamma2) update the 100-d memory:
    mem_t = gamma1 * mem_{t-1} + gamma2 * cHat.  (out_fc1 / out_fc2 are constructed but never used.)
    ``masks``: None (eval: the four Dropout(0.2) layers are identities) or four pre-scaled float masks (T, n, 100)
    multiplying the ReLU outputs of att1_fc1, att2_fc1, gamma1_fc1, gamma2_fc1 (:92,94,96,97)."""
    pre = (prefix + ".") if prefix else ""
    lin = lambda name, z: linear(z, P[f"{pre}{name}.weight"], P[f"{pre}{name}.bias"])
    T, n = x.shape[0], x.shape[1]
    xs = {"l": x[:, :, :300], "a": x[:, :, 300:600], "v": x[:, :, 600:]}
    h = {m: x.new_zeros(n, 100) for m in "lav"}
    c = {m: x.new_zeros(n, 100) for m in "lav"}
    mem = x.new_zeros(n, 100)
    outs = []
    drop = (lambda i, t, z: z * masks[i][t]) if masks is not None else (lambda i, t, z: z)
    for t in range(T):
        prev_cs = torch.cat([c["l"], c["a"], c["v"]], dim=1)
        for m in "lav":
            h[m], c[m] = mfn_lstm_cell(xs[m][t], h[m], c[m], P, f"{pre}lstm_{m}")
        c_star = torch.cat([prev_cs, c["l"], c["a"], c["v"]], dim=1)
        attention = torch.softmax(lin("att1_fc2", drop(0, t, torch.relu(lin("att1_fc1", c_star)))), dim=1)
        attended = attention * c_star
        c_hat = torch.tanh(lin("att2_fc2", drop(1, t, torch.relu(lin("att2_fc1", attended)))))
        both = torch.cat([attended, mem], dim=1)
        gamma1 = torch.sigmoid(lin("gamma1_fc2", drop(2, t, torch.relu(lin("gamma1_fc1", both)))))
        gamma2 = torch.sigmoid(lin("gamma2_fc2", drop(3, t, torch.relu(lin("gamma2_fc1", both)))))
        mem = gamma1 * mem + gamma2 * c_hat
        outs.append(torch.cat([h["l"], h["a"], h["v"], mem], dim=-1))
    return torch.stack(outs)


def mfn_head(feat: Tensor, lengths: Sequence[int], P: Dict[str, Tensor], masks: Optional[dict] = None) -> Tensor:
    """att_type 'mfn' after the graph model (code/model.py:1303-1330): node features (N, 900) padded per dialogue to the
    time-major window (T, B, 900), MFN, valid rows back in node order, dropout -> ReLU -> smax_fc (400 -> C) ->
    log_softmax.  masks: {'mfn': four pre-scaled (T, B, 100) masks, 'mfn_head': pre-scaled (N, 400)}."""
    mk = masks or {}
    T, B = int(max(lengths)), len(lengths)
    x = feat.new_zeros(T, B, feat.shape[1])
    off = 0
    for b, L in enumerate(lengths):
        x[:L, b] = feat[off:off + L]
        off += L
    out = mfn_forward(x, P, "mfn", mk.get("mfn"))
    rows = torch.cat([out[:L, b] for b, L in enumerate(lengths)], dim=0)
    if mk.get("mfn_head") is not None:
        rows = rows * mk["mfn_head"]
    return torch.log_softmax(linear(torch.relu(rows), P["smax_fc.weight"], P["smax_fc.bias"]), 1)


# --------------------------------------------------------------------------------------
# f4 (☆): low-rank multimodal fusion, code/model_fusion.py:214-310 (att_type='lmf_only')
# --------------------------------------------------------------------------------------
def lmf_forward(xa: Tensor, xv: Tensor, xt: Tensor, P: Dict[str, Tensor], prefix: str = "") -> Tensor:
    """LMF.forward (:275-310): per modality h = Linear(x); fusion_m = [1, h] . factor_m (rank, 301, 300) -> (rank, N, 300);
    out = fusion_weights (1, rank) . (fusion_a * fusion_v * fusion_t) + fusion_bias."""
    fz = None
    for x, n in ((xa, "audio"), (xv, "video"), (xt, "text")):
        h = linear(x, P[f"{prefix}{n}_subnet.weight"], P[f"{prefix}{n}_subnet.bias"])
        h1 = torch.cat([torch.ones(h.shape[0], 1, dtype=h.dtype), h], dim=1)
        f = torch.matmul(h1, P[f"{prefix}{n}_factor"])                       # (rank, N, 300)
        fz = f if fz is None else fz * f
    out = torch.matmul(P[prefix + "fusion_weights"], fz.permute(1, 0, 2)).squeeze(1) + P[prefix + "fusion_bias"]
    return out.view(-1, P[prefix + "fusion_bias"].shape[1])


# --------------------------------------------------------------------------------------
# f3 (☆): nodal-attention head of the `relation` graph type (text-only DialogueGCN configuration)
# --------------------------------------------------------------------------------------
def nodal_attention(E: Tensor, lengths: Sequence[int], w_t: Tensor, b_t: Tensor) -> Tensor:
    """attentive_node_features (code/model.py:614-645) with MatchingAttention('general2') (:66-76, :83-84), restated on
    the ragged rows.  For dialogue b with rows E_b (L, D): x_ = transform(x_t) = W x_t + b; alpha_ = tanh(x_ . M_s) on the
    valid positions (0 on the padded ones), softmax over ALL positions, re-masked and re-normalised -- the padded
    positions' exp(0) cancels, leaving a softmax of tanh(scores) over the valid positions; pool = sum_s alpha_s M_s.
    (The reference also evaluates padded candidates t >= L; classify_node_features drops them, :663.)"""
    out, off = [], 0
    for L in lengths:
        Eb = E[off:off + L]
        Q = Eb @ w_t.t() + b_t
        P = torch.softmax(torch.tanh(Q @ Eb.t()), dim=1)
        out.append(P @ Eb)
        off += L
    return torch.cat(out, 0)


def nodal_head(E: Tensor, lengths: Sequence[int], P: Dict[str, Tensor], mask: Optional[Tensor] = None, scale: float = 1.0,
               prefix: str = "") -> Tensor:
    """classify_node_features(nodal_attn=True, avec=False) (code/model.py:647-664): attentive features ->
    relu(linear) -> dropout (`mask`: (N, hidden) keep mask, None = identity) -> smax_fc -> log_softmax."""
    att = nodal_attention(E, lengths, P[prefix + "matchatt.transform.weight"], P[prefix + "matchatt.transform.bias"])
    hid = torch.relu(att @ P[prefix + "linear.weight"].t() + P[prefix + "linear.bias"])
    if mask is not None:
        hid = hid * mask.to(hid.dtype) * scale
    return torch.log_softmax(hid @ P[prefix + "smax_fc.weight"].t() + P[prefix + "smax_fc.bias"], dim=1)


# --------------------------------------------------------------------------------------
# deterministic, torch-version-independent weights and synthetic inputs (shared by the
# golden generator, the tests and bench.py so that nothing has to travel to the GPU box)
# --------------------------------------------------------------------------------------
def formula_weights(shapes: Dict[str, Tuple[int, ...]], seed: int = 2021) -> Dict[str, Tensor]:
    """U(-s, s) per tensor, s = 1/sqrt(last dim) for matrices, 0.05 for vectors; one
    ``np.random.RandomState(seed + k)`` per key, k = rank of the key in sorted order."""
    out = {}
    for k, name in enumerate(sorted(shapes)):
        shp = tuple(shapes[name])
        rs = np.random.RandomState(seed + k)
        s = 1.0 / math.sqrt(shp[-1]) if len(shp) >= 2 else 0.05
        out[name] = torch.from_numpy(rs.uniform(-s, s, size=shp).astype(np.float32))
    return out


def synthetic_batch(lengths: Sequence[int], d_text: int, d_audio: int, d_visual: int, n_speakers: int,
                    n_classes: int, seed: int = 0, T: Optional[int] = None):
    """Seeded synthetic batch in the reference's collate layout (code/dataloader.py:31-34):
    time-major zero-padded features, one-hot qmask (zero on padding), umask, packed labels."""
    rs = np.random.RandomState(seed)
    B = len(lengths)
    T = int(max(lengths)) if T is None else T
    def feat(d):
        x = rs.standard_normal((T, B, d)).astype(np.float32)
        for b, L in enumerate(lengths):
            x[L:, b] = 0
        return torch.from_numpy(x)
    textf, acouf, visuf = feat(d_text), feat(d_audio), feat(d_visual)
    spk = rs.randint(0, n_speakers, size=(T, B))
    qmask = np.zeros((T, B, n_speakers), np.float32)
    umask = np.zeros((B, T), np.float32)
    for b, L in enumerate(lengths):
        qmask[np.arange(L), b, spk[:L, b]] = 1
        umask[b, :L] = 1
    label = torch.from_numpy(np.concatenate([rs.randint(0, n_classes, size=L) for L in lengths]).astype(np.int64))
    return textf, acouf, visuf, torch.from_numpy(qmask), torch.from_numpy(umask), label
