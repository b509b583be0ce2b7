"""f4: GCNII (code/model_GCN.py:224-306) and graph_type='DeepGCN' (code/model.py:922-940, 1244-1293) on the GPU path:
the stand-alone module against the UNMODIFIED reference class (tests/golden/gcnii.npz: with the fusion gate forward only --
the reference's own backward raises there --, without it forward + input / parameter gradients), against the oracle with
injected dropout masks, and the DeepGCN model (three GCNII networks with their own weights on the shared zero-cross-weight
adjacency -> concat / gated fusion -> head) against the oracle's composition."""
import os

import numpy as np
import pytest
import torch

import mmdfn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def _net(K, reason, dropout=0.0):
    from mmdfn_b200.modules import GCNII
    return GCNII(nfeat=200, nlayers=K, nhidden=100, nclass=6, dropout=dropout, lamda=0.5, alpha=0.1, variant=True,
                 return_feature=True, use_residue=True, reason_flag=reason)


@pytest.mark.parametrize("tag,K,reason", [("r", 3, True), ("n", 4, False)])
def test_matches_reference_golden(tag, K, reason):
    g = np.load(os.path.join(HERE, "golden", "gcnii.npz"))
    lengths = [int(x) for x in g["lengths"]]
    net = _net(K, reason)
    assert sorted(net.state_dict().keys()) == list(g[tag + ".keys"])
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(O.formula_weights(shapes, seed=61 + K), strict=True)
    net = net.to(DEV).train()
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    out = net(x, lengths, None)
    assert float((out.detach().cpu() - torch.from_numpy(g[tag + ".out"])).abs().max()) < 1e-4
    if reason:
        return
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    r = torch.from_numpy(g["n.dx"])
    assert float((x.grad.cpu() - r).norm() / float(r.norm())) < 1e-3
    for k, p in net.named_parameters():
        if not bool(g["n.used." + k]):
            assert p.grad is None, k
            continue
        ref = float(g["n.gnorm." + k])
        assert abs(float(p.grad.norm()) - ref) < 1e-3 * max(1.0, ref), k


def test_masks_and_gate_gradients_match_oracle():
    lengths, K = [14, 5, 22], 3
    N = sum(lengths)
    net = _net(K, True, dropout=0.5)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    P0 = O.formula_weights(shapes, seed=5)
    net.load_state_dict(P0, strict=True)
    rs = np.random.RandomState(4)
    x0 = torch.from_numpy((0.7 * rs.standard_normal((N, 200))).astype(np.float32))
    G = torch.from_numpy(rs.standard_normal((N, 300)).astype(np.float32))
    mk = {"x": torch.from_numpy((rs.rand(N, 200) > 0.5).astype(np.uint8)), "h0": torch.from_numpy((rs.rand(N, 100) > 0.5).astype(np.uint8)),
          "out": torch.from_numpy((rs.rand(N, 100) > 0.5).astype(np.uint8))}
    P = {"net." + k: v.clone().requires_grad_(True) for k, v in P0.items()}
    xr = x0.clone().requires_grad_(True)
    ref = O.gcnii(xr, lengths, P, "net", K, 0.5, 0.1, reason_flag=True, masks={k: v.float() * 2.0 for k, v in mk.items()})
    (ref * G).sum().backward()
    net = net.to(DEV).train()
    x = x0.to(DEV).requires_grad_(True)
    out = net(x, lengths, None, masks={k: v.to(DEV) for k, v in mk.items()})
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 1e-4
    (out * G.to(DEV)).sum().backward()
    assert float((x.grad.cpu() - xr.grad).norm() / float(xr.grad.norm())) < 2e-3
    for k, p in net.named_parameters():
        r = P["net." + k].grad
        assert float((p.grad.cpu() - r).norm() / max(float(r.norm()), 1e-8)) < 2e-3, k


@pytest.mark.parametrize("att,reason", [("concat_subsequently", True), ("gated", False)])
def test_deep_gcn_model_vs_oracle(att, reason):
    import mmdfn_b200 as mm
    from mmdfn_b200.dp import used_parameters
    lengths, S, C, K = [9, 14, 6], 2, 6, 2
    t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 40, 24, S, C, seed=43)
    m = mm.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.0, graph_type="DeepGCN", D_m_v=24, D_m_a=40, modals="avl",
                            att_type=att, Deep_GCN_nlayers=K, use_speaker=False, reason_flag=reason, use_crn_speaker=True,
                            speaker_weights="1-0.5-2")
    m.load_state_dict(O.formula_weights({k: tuple(p.shape) for k, p in m.state_dict().items()}, seed=59))
    P = {k: p.detach().clone().requires_grad_(True) for k, p in m.state_dict().items()}
    wts = (1.0, 0.5, 2.0)
    U_a, U_v, U_l = (O.linear(x, P[f"linear_{n}.weight"], P[f"linear_{n}.bias"]) for x, n in ((a, "a"), (v, "v"), (t, "l")))
    E_l = O.bigru2(U_l, P, "lstm_l")
    em = [U_a + wts[0] * O.party_encode(U_a, q, P), U_v + wts[1] * O.party_encode(U_v, q, P), E_l + wts[2] * O.party_encode(U_l, q, P)]
    e = [O.gcnii(O.ragged_pack(x, lengths), lengths, P, "graph_net_" + n, K, 0.5, 0.1, reason_flag=reason) for x, n in zip(em, "avl")]
    feat = torch.cat(e, -1) if att == "concat_subsequently" else O.mm_gated_attention(e[0], e[1], e[2], P)
    lp_ref = torch.log_softmax(O.linear(torch.relu(feat), P["smax_fc.weight"], P["smax_fc.bias"]), 1)
    O.focal_loss(lp_ref, lab, 1.0).backward()
    m = m.to(DEV).train()
    m.gatedatt.eval()
    lp = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV))[0]
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 1e-4
    mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV)).backward()
    for k, p in m.named_parameters():
        if P[k].grad is None:
            assert p.grad is None, k
            continue
        g, r = p.grad.cpu(), P[k].grad
        assert float((g - r).norm() / max(float(r.norm()), 1e-8)) < 2e-3, k
    assert {n for n, _ in used_parameters(m)} == {k for k, p in m.named_parameters() if p.grad is not None}
