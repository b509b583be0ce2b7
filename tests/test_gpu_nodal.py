"""f3: the nodal-attention classifier head of the `relation` graph type on the GPU path --
`classify_node_features` / `attentive_node_features` / `MatchingAttention('general2')` (code/model.py:614-672, 66-76) --
against (1) the outputs of the UNMODIFIED reference functions stored in tests/golden/nodal_head.npz (log-probabilities,
attentive features, input gradient, parameter-gradient summaries) and (2) the oracle's ragged restatement with autograd
gradients on larger ragged batches incl. an injected dropout mask and dialogue lengths beyond one 32-row tile.
Tolerances: 2e-5 on log-probabilities / features (fp32 tanh, exp, FFMA re-association), 2e-4 relative on gradients."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import mmdfn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def _layers(P, D, hid, C, p_drop):
    import mmdfn_b200
    from mmdfn_b200.modules import MatchingAttention
    m = {"matchatt": MatchingAttention(D, D, att_type='general2'), "linear": nn.Linear(D, hid), "smax_fc": nn.Linear(hid, C)}
    for n, mod in m.items():
        mod.load_state_dict({k: P[f"{n}.{k}"] for k in mod.state_dict()}, strict=True)
        mod.to(DEV)
    return m, nn.Dropout(p_drop)


def test_matches_reference_golden():
    from mmdfn_b200 import relation as R
    g = np.load(os.path.join(HERE, "golden", "nodal_head.npz"))
    lengths = [int(x) for x in g["lengths"]]
    P = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}
    m, drop = _layers(P, 300, 100, 6, 0.0)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    umask = torch.zeros(len(lengths), max(lengths), device=DEV)
    for b, L in enumerate(lengths):
        umask[b, :L] = 1
    lp = R.classify_node_features(x, lengths, umask, m["matchatt"], m["linear"], drop, m["smax_fc"], True, False, False)
    assert lp.shape == g["log_prob"].shape
    assert float((lp.detach().cpu() - torch.from_numpy(g["log_prob"])).abs().max()) < 2e-5
    att = R.attentive_node_features(x.detach(), lengths, umask, m["matchatt"], False).cpu()
    off = 0
    for b, L in enumerate(lengths):
        assert float((att[off:off + L] - torch.from_numpy(g["att_padded"][:L, b])).abs().max()) < 2e-5
        off += L
    (lp * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    assert float((x.grad.cpu() - torch.from_numpy(g["dx"])).abs().max()) < 2e-5 * max(1.0, float(np.abs(g["dx"]).max()))
    for n, mod in m.items():
        for k, p in mod.named_parameters():
            ref = float(g[f"gnorm.{n}.{k}"])
            assert abs(float(p.grad.norm()) - ref) < 2e-4 * max(1.0, ref), (n, k)
            assert abs(float(p.grad.sum()) - float(g[f"gsum.{n}.{k}"])) < 2e-4 * max(1.0, ref), (n, k)


@pytest.mark.parametrize("lengths,D,with_mask", [([1], 300, False), ([33, 2, 64, 17], 300, True), ([110, 59, 91, 8, 31], 300, True),
                                                 ([40, 70, 5], 400, False), ([200, 129], 300, False)])
def test_forward_and_gradients_match_oracle(lengths, D, with_mask):
    from mmdfn_b200 import relation as R
    hid, C = 100, 7
    shapes = {"matchatt.transform.weight": (D, D), "matchatt.transform.bias": (D,), "linear.weight": (hid, D), "linear.bias": (hid,),
              "smax_fc.weight": (C, hid), "smax_fc.bias": (C,)}
    P = O.formula_weights(shapes, seed=11)
    m, drop = _layers(P, D, hid, C, 0.5)
    rs = np.random.RandomState(sum(lengths))
    N = sum(lengths)
    x0 = torch.from_numpy((0.4 * rs.standard_normal((N, D))).astype(np.float32))
    G = torch.from_numpy(rs.standard_normal((N, C)).astype(np.float32))
    mask = torch.from_numpy((rs.rand(N, hid) > 0.5).astype(np.uint8)) if with_mask else None
    # oracle (CPU fp32, autograd)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xr = x0.clone().requires_grad_(True)
    lp_ref = O.nodal_head(xr, lengths, Pr, mask=mask, scale=2.0)
    (lp_ref * G).sum().backward()
    # GPU path
    x = x0.to(DEV).requires_grad_(True)
    drop.eval() if not with_mask else drop.train()
    lp = R.classify_node_features(x, lengths, None, m["matchatt"], m["linear"], drop, m["smax_fc"], True, False, False,
                                  mask=mask.to(DEV) if with_mask else None)
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 2e-5
    (lp * G.to(DEV)).sum().backward()
    gx = xr.grad
    assert float((x.grad.cpu() - gx).norm() / max(float(gx.norm()), 1e-12)) < 2e-4
    for n, mod in m.items():
        for k, p in mod.named_parameters():
            ref = Pr[f"{n}.{k}"].grad
            assert float((p.grad.cpu() - ref).norm() / max(float(ref.norm()), 1e-12)) < 2e-4, (n, k)


def test_graph_network_with_nodal_head_runs_and_normalises():
    """GraphNetwork(return_feature=False): RGCN -> GraphConv -> nodal-attention head on the relation path's edge set"""
    from mmdfn_b200 import relation as R
    torch.manual_seed(3)
    lengths, S, Dn = [12, 5, 20], 2, 200
    N, T, B = sum(lengths), max(lengths), len(lengths)
    net = R.GraphNetwork(Dn, 6, 2 * S * S, 200, hidden_size=100, dropout=0.5, no_cuda=False, return_feature=False).to(DEV).eval()
    rs = np.random.RandomState(5)
    feats = torch.from_numpy(rs.standard_normal((T, B, Dn)).astype(np.float32)).to(DEV)
    qmask = torch.zeros(T, B, S, device=DEV)
    spk = torch.from_numpy(rs.randint(0, S, size=(T, B))).to(DEV)
    qmask.scatter_(2, spk.unsqueeze(-1), 1.0)
    umask = torch.zeros(B, T, device=DEV)
    for b, L in enumerate(lengths):
        umask[b, :L] = 1
    import mmdfn_b200
    model_att = mmdfn_b200.MaskedEdgeAttention(Dn, 200, False).to(DEV)
    mapping = {}
    for j in range(S):
        for k in range(S):
            mapping[str(j) + str(k) + '0'] = len(mapping)
            mapping[str(j) + str(k) + '1'] = len(mapping)
    x, edge_index, edge_norm, edge_type, ell = R.batch_graphify(feats, qmask, lengths, 10, 10, mapping, model_att, False)
    x = x.detach().requires_grad_(True)
    lp = net(x, edge_index, edge_norm, edge_type, lengths, umask, True, False)
    assert lp.shape == (N, 6)
    assert float((lp.exp().sum(1) - 1).abs().max()) < 1e-5
    lp.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("modals,hidden_,nodal", [("l", 100, True), ("avl", 250, True), ("l", 100, False)])
def test_single_stream_relation_model_end_to_end(modals, hidden_, nodal):
    """DialogueGNNModel(multi_modal=False, graph_type='relation') -- the DialogueGCN configuration (code/model.py:828-849,
    1035-1036, 1176-1180, 1211-1212): linear_ -> BiGRU (layer-0 input width 100 / 250) -> windowed relation graph -> RGCN ->
    GraphConv -> nodal-attention head; log-probabilities, edges and every parameter gradient vs the oracle's composition
    (inter-layer GRU dropout and head dropout masks injected)."""
    import mmdfn_b200 as mm
    lengths, S, C, D_m = [13, 6, 21, 9], 2, 6, 64
    N, T, B = sum(lengths), max(lengths), len(lengths)
    t, a, v, q, u, lab = O.synthetic_batch(lengths, D_m, 8, 8, S, C, seed=31)
    m = mm.DialogueGNNModel("LSTM", D_m, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.5, nodal_attention=nodal, graph_type="relation", modals=modals,
                            att_type="concat", multi_modal=False, use_crn_speaker=False)
    assert tuple(m.linear_.weight.shape) == (hidden_, D_m) and tuple(m.lstm.weight_ih_l0.shape) == (300, hidden_)
    m.load_state_dict(O.formula_weights({k: tuple(p.shape) for k, p in m.state_dict().items()}, seed=41))
    P = {k: p.detach().clone().requires_grad_(True) for k, p in m.state_dict().items()}
    rs = np.random.RandomState(9)
    m_gru = torch.from_numpy((rs.rand(T, B, 200) > 0.5).astype(np.uint8))
    m_head = torch.from_numpy((rs.rand(N, 100) > 0.5).astype(np.uint8))
    # ---- oracle
    x = O.linear(t, P["linear_.weight"], P["linear_.bias"])
    em = O.bigru2(x, P, "lstm", inter_mask=m_gru.float() * 2.0)
    ei, et, counts = O.build_edges(q.numpy(), lengths, 10, 10)
    en = O.edge_norms(O.masked_edge_attention(em, P["att_model.scalar.weight"], lengths, 10, 10), lengths, 10, 10)
    xr = O.ragged_pack(em, lengths)
    h1 = O.rgcn_conv(xr, ei, et, en, P["graph_net.conv1.basis"], P["graph_net.conv1.att"], P["graph_net.conv1.root"], P["graph_net.conv1.bias"])
    h2 = O.pyg_graph_conv(h1, ei, P["graph_net.conv2.weight"], P["graph_net.conv2.lin.weight"], P["graph_net.conv2.lin.bias"])
    feat = torch.cat([xr, h2], -1)
    if nodal:
        lp_ref = O.nodal_head(feat, lengths, P, mask=m_head, scale=2.0, prefix="graph_net.")
    else:
        hid = torch.relu(O.linear(feat, P["graph_net.linear.weight"], P["graph_net.linear.bias"])) * m_head.float() * 2.0
        lp_ref = torch.log_softmax(O.linear(hid, P["graph_net.smax_fc.weight"], P["graph_net.smax_fc.bias"]), 1)
    O.focal_loss(lp_ref, lab, 1.0).backward()
    # ---- kernels
    m = m.to(DEV).train()
    lp, edge_index, edge_norm, edge_type, eil = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths,
                                                  masks={"gru": m_gru.to(DEV), "head": m_head.to(DEV)})
    assert np.array_equal(edge_index.cpu().numpy(), ei) and np.array_equal(edge_type.cpu().numpy(), et) and eil == counts
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 1e-4
    mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV)).backward()
    for k, p in m.named_parameters():
        if P[k].grad is None:
            assert p.grad is None, k
            continue
        g, r = p.grad.cpu(), P[k].grad
        assert float((g - r).norm() / max(float(r.norm()), 1e-8)) < 1e-3, k
    # the trainer's parameter selection for this configuration covers exactly the parameters that received a gradient
    from mmdfn_b200.dp import used_parameters
    used = {n for n, _ in used_parameters(m)}
    assert used == {k for k, p in m.named_parameters() if p.grad is not None}
