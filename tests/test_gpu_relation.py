"""GPU parity of the `relation`-path kernels: windowed edges (bit-exact on the sorted
list), edge types, masked edge attention forward/backward -- against the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import mmdfn_oracle as O
from helpers import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mods():
    import mmdfn_b200
    from mmdfn_b200 import ops, relation
    return mmdfn_b200, ops, relation


@pytest.mark.parametrize("tag,S", [("rel2", 2), ("rel9", 9)])
def test_batch_graphify_matches_reference_golden(tag, S):
    mm, ops, rel = _mods()
    s = load_case("submodules")
    lens = [int(x) for x in s[f"{tag}_lens"]]
    feats = torch.from_numpy(s[f"{tag}_feats"]).to(DEV)
    qmask = torch.from_numpy(s[f"{tag}_qmask"]).to(DEV)
    att = mm.MaskedEdgeAttention(200, 200, False)
    att.load_state_dict(O.formula_weights({k: tuple(t.shape) for k, t in att.state_dict().items()}, seed=9))
    att = att.to(DEV)
    assert np.array_equal(att.scalar.weight.detach().cpu().numpy(), s[f"{tag}_att_w"])
    with torch.no_grad():
        nf, ei, en, et, eil = rel.batch_graphify(feats, qmask, lens, 10, 10, {}, att, False)
    assert np.array_equal(ei.cpu().numpy(), s[f"{tag}_edge_index"])            # bit-exact (sorted edge list)
    assert np.array_equal(et.cpu().numpy(), s[f"{tag}_edge_type"])
    assert eil == [int(x) for x in s[f"{tag}_edge_lens"]]
    assert float(np.abs(en.cpu().numpy() - s[f"{tag}_edge_norm"]).max()) < 1e-6
    assert torch.equal(nf.cpu(), torch.from_numpy(s[f"{tag}_node_features"]))


@pytest.mark.parametrize("lengths,S,wp,wf", [([1], 2, 10, 10), ([37, 12, 50], 2, 10, 10), ([5, 1, 23], 9, 3, 7),
                                             ([20, 60], 2, -1, -1), ([64, 3], 3, -1, 4), ([30, 31], 2, 0, 0)])
def test_edges_bit_exact_and_attention_grads(lengths, S, wp, wf):
    mm, ops, rel = _mods()
    rs = np.random.RandomState(len(lengths) * 7 + S)
    T, B = max(lengths), len(lengths)
    q = np.zeros((T, B, S), np.float32)
    spk = rs.randint(0, S, size=(T, B))
    for b, L in enumerate(lengths):
        q[np.arange(L), b, spk[:L, b]] = 1
    ei_ref, et_ref, counts = O.build_edges(q, lengths, wp, wf)
    geom = ops.DialogGeom(lengths, DEV)
    edges = rel.EdgeSet(torch.from_numpy(q).to(DEV), geom, wp, wf)
    assert edges.counts == counts and edges.E == ei_ref.shape[1]
    assert np.array_equal(edges.edge_index.cpu().numpy(), ei_ref)
    assert np.array_equal(edges.edge_type.cpu().numpy(), et_ref)
    # attention forward / backward against autograd over the oracle
    M = torch.from_numpy(rs.standard_normal((T, B, 200)).astype(np.float32))
    W = torch.from_numpy((rs.standard_normal((200, 200)) * 0.07).astype(np.float32))
    Mc, Wc = M.clone().requires_grad_(True), W.clone().requires_grad_(True)
    sc_ref = O.masked_edge_attention(Mc, Wc, lengths, wp, wf)
    en_ref = O.edge_norms(sc_ref, lengths, wp, wf)
    g = torch.from_numpy(rs.standard_normal(en_ref.shape[0]).astype(np.float32))
    (en_ref * g).sum().backward()
    Mg, Wg = M.to(DEV).requires_grad_(True), W.to(DEV).requires_grad_(True)
    en = rel.EdgeAttnFn.apply(Mg, Wg, edges)
    assert float((en.detach().cpu() - en_ref.detach()).abs().max()) < 1e-6
    dense = rel.ScoresDenseFn.apply(en, edges, 200, T)
    assert float((dense.detach().cpu() - sc_ref.detach()).abs().max()) < 1e-6
    if wp == 0 and wf == 0:
        return          # self-loops only: edge_norm == 1 up to 1e-10, the true gradient is ~1e-10 rounding noise
    (en * g.to(DEV)).sum().backward()
    rel_err = lambda a, b: float((a.cpu() - b).norm() / max(float(b.norm()), 1e-4))   # wp=wf=0: true gradient ~1e-10
    assert rel_err(Mg.grad, Mc.grad) < 1e-4
    assert rel_err(Wg.grad, Wc.grad) < 1e-4


def test_masked_edge_attention_module_api():
    mm, ops, rel = _mods()
    lengths = [20, 33]
    T, B = 33, 2
    M = torch.randn(T, B, 200, generator=torch.Generator().manual_seed(0))
    att = mm.MaskedEdgeAttention(200, 200, False)
    att.load_state_dict(O.formula_weights({k: tuple(t.shape) for k, t in att.state_dict().items()}, seed=4))
    edge_ind = [rel.edge_perms(L, 10, 10) for L in lengths]
    with torch.no_grad():
        sc = att.to(DEV)(M.to(DEV), lengths, edge_ind)
    ref = O.masked_edge_attention(M, att.scalar.weight.detach().cpu(), lengths, 10, 10)
    assert sc.shape == (B, 200, T)
    assert float((sc.cpu() - ref).abs().max()) < 1e-6


@pytest.mark.parametrize("lengths,S,wp,wf", [([37, 12, 50], 2, 10, 10), ([5, 1, 23], 9, 3, 7), ([30, 8], 3, -1, 4)])
def test_graph_network_rgcn_graphconv_vs_oracle(lengths, S, wp, wf):
    """a12: RGCNConv -> GraphConv (PyG 1.4.3 semantics as restated in the oracle; parity unpinned by the reference)."""
    mm, ops, rel = _mods()
    rs = np.random.RandomState(5 + S)
    T, B, N = max(lengths), len(lengths), sum(lengths)
    q = np.zeros((T, B, S), np.float32)
    spk = rs.randint(0, S, size=(T, B))
    for b, L in enumerate(lengths):
        q[np.arange(L), b, spk[:L, b]] = 1
    feats = torch.from_numpy(rs.standard_normal((T, B, 200)).astype(np.float32))
    att = mm.MaskedEdgeAttention(200, 200, False)
    att.load_state_dict(O.formula_weights({k: tuple(t.shape) for k, t in att.state_dict().items()}, seed=9))
    R = 2 * S * S
    net = rel.GraphNetwork(200, 6, R, 200, 100, 0.4, False, False, True)
    net.load_state_dict(O.formula_weights({k: tuple(t.shape) for k, t in net.state_dict().items()}, seed=13))
    P = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    Wc = att.scalar.weight.detach().clone().requires_grad_(True)
    # oracle
    fc = feats.clone().requires_grad_(True)
    ei, et, counts = O.build_edges(q, lengths, wp, wf)
    en = O.edge_norms(O.masked_edge_attention(fc, Wc, lengths, wp, wf), lengths, wp, wf)
    x = O.ragged_pack(fc, lengths)
    h1 = O.rgcn_conv(x, ei, et, en, P["conv1.basis"], P["conv1.att"], P["conv1.root"], P["conv1.bias"])
    h2 = O.pyg_graph_conv(h1, ei, P["conv2.weight"], P["conv2.lin.weight"], P["conv2.lin.bias"])
    out_ref = torch.cat([x, h2], -1)
    g = torch.from_numpy(rs.standard_normal(out_ref.shape).astype(np.float32))
    (out_ref * g).sum().backward()
    # kernels
    net, att = net.to(DEV), att.to(DEV)
    fg = feats.to(DEV).requires_grad_(True)
    nf, edge_index, edge_norm, edge_type, eil = rel.batch_graphify(fg, torch.from_numpy(q).to(DEV), lengths, wp, wf, {}, att, False)
    out = net(nf, edge_index, edge_norm, edge_type, lengths, None, False, False)
    assert out.shape == (N, 300)
    err = float((out.detach().cpu() - out_ref.detach()).abs().max())
    assert err < 2e-5, err
    (out * g.to(DEV)).sum().backward()
    rel_err = lambda a, b: float((a.cpu() - b).norm() / max(float(b.norm()), 1e-8))
    assert rel_err(fg.grad, fc.grad) < 2e-4
    assert rel_err(att.scalar.weight.grad, Wc.grad) < 2e-4
    for k, p in net.named_parameters():
        assert rel_err(p.grad, P[k].grad) < 2e-4, k


@pytest.mark.parametrize("lengths,S", [([9, 14, 6], 2), ([12, 1, 7, 20], 4)])
def test_dialogue_gnn_model_relation_graph_type(lengths, S):
    """graph_type='relation' end to end (encoders -> windowed edges -> edge attention -> RGCN/GraphConv per modality ->
    head without ReLU) against the oracle composition of code/model.py:1182-1242; logits and all gradients."""
    mm, ops, rel = _mods()
    C, dT, dA, dV = 6, 100, 40, 24
    t, a, v, q, u, lab = O.synthetic_batch(lengths, dT, dA, dV, S, C, seed=21)
    m = mm.DialogueGNNModel("LSTM", dT, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.0, graph_type="relation", alpha=0.2, lamda=0.5,
                            D_m_v=dV, D_m_a=dA, modals="avl", att_type="concat_subsequently", Deep_GCN_nlayers=2,
                            use_speaker=False, reason_flag=False, use_crn_speaker=True, speaker_weights="1-0.5-2")
    m.load_state_dict(O.formula_weights({k: tuple(p.shape) for k, p in m.state_dict().items()}, seed=77))
    P = {k: p.detach().clone().requires_grad_(True) for k, p in m.state_dict().items()}
    # ---- oracle composition
    wts = (1.0, 0.5, 2.0)
    U_a, U_v, U_l = (O.linear(x, P[f"linear_{n}.weight"], P[f"linear_{n}.bias"]) for x, n in ((a, "a"), (v, "v"), (t, "l")))
    E_l = O.bigru2(U_l, P, "lstm_l")
    em = [U_a + wts[0] * O.party_encode(U_a, q, P), U_v + wts[1] * O.party_encode(U_v, q, P), E_l + wts[2] * O.party_encode(U_l, q, P)]
    ei, et, counts = O.build_edges(q.numpy(), lengths, 10, 10)
    en = O.edge_norms(O.masked_edge_attention(em[2], P["att_model.scalar.weight"], lengths, 10, 10), lengths, 10, 10)
    outs = []
    for e, n in zip(em, "avl"):
        x = O.ragged_pack(e, lengths)
        h1 = O.rgcn_conv(x, ei, et, en, P[f"graph_net_{n}.conv1.basis"], P[f"graph_net_{n}.conv1.att"],
                         P[f"graph_net_{n}.conv1.root"], P[f"graph_net_{n}.conv1.bias"])
        h2 = O.pyg_graph_conv(h1, ei, P[f"graph_net_{n}.conv2.weight"], P[f"graph_net_{n}.conv2.lin.weight"],
                              P[f"graph_net_{n}.conv2.lin.bias"])
        outs.append(torch.cat([x, h2], -1))
    lp_ref = torch.log_softmax(O.linear(torch.cat(outs, -1), P["smax_fc.weight"], P["smax_fc.bias"]), 1)
    loss_ref = O.focal_loss(lp_ref, lab, 1.0)
    loss_ref.backward()
    # ---- kernels
    m = m.to(DEV).train()
    lp, edge_index, edge_norm, edge_type, eil = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV))
    assert np.array_equal(edge_index.cpu().numpy(), ei) and np.array_equal(edge_type.cpu().numpy(), et) and eil == counts
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 1e-4
    loss = mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV))
    loss.backward()
    for k, p in m.named_parameters():
        if P[k].grad is None:
            assert p.grad is None, k
            continue
        g, r = p.grad.cpu(), P[k].grad
        assert float((g - r).norm() / max(float(r.norm()), 1e-8)) < 1e-3, k


def test_dialogue_gnn_model_relation_gated_attention():
    """graph_type='relation', att_type='gated' (code/model.py:1235-1239): MMGatedAttention over the three networks'
    features (its Dropout(0.5) masks injected) -> smax_fc (300 -> C) -> log_softmax; logits and all gradients vs the oracle."""
    mm, ops, rel = _mods()
    lengths, S, C, dT, dA, dV = [9, 14, 6], 2, 6, 100, 40, 24
    N = sum(lengths)
    t, a, v, q, u, lab = O.synthetic_batch(lengths, dT, dA, dV, S, C, seed=23)
    m = mm.DialogueGNNModel("LSTM", dT, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.0, graph_type="relation", alpha=0.2, lamda=0.5,
                            D_m_v=dV, D_m_a=dA, modals="avl", att_type="gated", Deep_GCN_nlayers=2,
                            use_speaker=False, reason_flag=False, use_crn_speaker=True, speaker_weights="1-0.5-2")
    assert tuple(m.smax_fc.weight.shape) == (C, 300)
    m.load_state_dict(O.formula_weights({k: tuple(p.shape) for k, p in m.state_dict().items()}, seed=78))
    P = {k: p.detach().clone().requires_grad_(True) for k, p in m.state_dict().items()}
    rs = np.random.RandomState(5)
    gmasks = [torch.from_numpy((rs.rand(N, 300) < 0.5).astype(np.uint8)) for _ in range(3)]
    wts = (1.0, 0.5, 2.0)
    U_a, U_v, U_l = (O.linear(x, P[f"linear_{n}.weight"], P[f"linear_{n}.bias"]) for x, n in ((a, "a"), (v, "v"), (t, "l")))
    E_l = O.bigru2(U_l, P, "lstm_l")
    em = [U_a + wts[0] * O.party_encode(U_a, q, P), U_v + wts[1] * O.party_encode(U_v, q, P), E_l + wts[2] * O.party_encode(U_l, q, P)]
    ei, et, counts = O.build_edges(q.numpy(), lengths, 10, 10)
    en = O.edge_norms(O.masked_edge_attention(em[2], P["att_model.scalar.weight"], lengths, 10, 10), lengths, 10, 10)
    outs = []
    for e, n in zip(em, "avl"):
        x = O.ragged_pack(e, lengths)
        h1 = O.rgcn_conv(x, ei, et, en, P[f"graph_net_{n}.conv1.basis"], P[f"graph_net_{n}.conv1.att"],
                         P[f"graph_net_{n}.conv1.root"], P[f"graph_net_{n}.conv1.bias"])
        h2 = O.pyg_graph_conv(h1, ei, P[f"graph_net_{n}.conv2.weight"], P[f"graph_net_{n}.conv2.lin.weight"],
                              P[f"graph_net_{n}.conv2.lin.bias"])
        outs.append(torch.cat([x, h2], -1))
    feat = O.mm_gated_attention(*[o * mk.float() * 2.0 for o, mk in zip(outs, gmasks)], P)
    lp_ref = torch.log_softmax(O.linear(feat, P["smax_fc.weight"], P["smax_fc.bias"]), 1)
    loss_ref = O.focal_loss(lp_ref, lab, 1.0)
    loss_ref.backward()
    m = m.to(DEV).train()
    lp = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV),
           masks={"gated": {"in": [mk.to(DEV) for mk in gmasks]}})[0]
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 1e-4
    mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV)).backward()
    for k, p in m.named_parameters():
        if P[k].grad is None:
            assert p.grad is None, k
            continue
        g, r = p.grad.cpu(), P[k].grad
        assert float((g - r).norm() / max(float(r.norm()), 1e-8)) < 1e-3, k
