"""f4: the graph-free multimodal baselines -- DialogueGNNModel(graph_type='None') with att_type concat_only / lmf_only /
mfn_only / gated (code/model.py:952-961, 1338-1405): encoders -> [Linear(200 -> 100)(x_m) | x_m] per modality -> fusion ->
dropout -> smax_fc -> log_softmax -- log-probabilities and every parameter gradient against the oracle's composition of the
reference-pinned pieces (encoders, MMGatedAttention, MFN, LMF), with the final dropout mask injected."""
import numpy as np
import pytest
import torch

import mmdfn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("att", ["concat_only", "lmf_only", "mfn_only", "gated"])
def test_no_graph_baselines_vs_oracle(att):
    import mmdfn_b200 as mm
    from mmdfn_b200.dp import used_parameters
    lengths, S, C, dT, dA, dV = [9, 14, 6, 11], 2, 6, 100, 40, 24
    N, T = sum(lengths), max(lengths)
    t, a, v, q, u, lab = O.synthetic_batch(lengths, dT, dA, dV, S, C, seed=37)
    m = mm.DialogueGNNModel("LSTM", dT, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.5, graph_type="None", D_m_v=dV, D_m_a=dA, modals="avl",
                            att_type=att, use_speaker=False, use_crn_speaker=True, speaker_weights="1-0.5-2")
    m.load_state_dict(O.formula_weights({k: tuple(p.shape) for k, p in m.state_dict().items()}, seed=53))
    P = {k: p.detach().clone().requires_grad_(True) for k, p in m.state_dict().items()}
    F = {"concat_only": 900, "lmf_only": 300, "mfn_only": 400, "gated": 300}[att]
    rs = np.random.RandomState(3)
    m_head = torch.from_numpy((rs.rand(N, F) > 0.5).astype(np.uint8))
    # ---- oracle
    wts = (1.0, 0.5, 2.0)
    U_a, U_v, U_l = (O.linear(x, P[f"linear_{n}.weight"], P[f"linear_{n}.bias"]) for x, n in ((a, "a"), (v, "v"), (t, "l")))
    E_l = O.bigru2(U_l, P, "lstm_l")
    em = [U_a + wts[0] * O.party_encode(U_a, q, P), U_v + wts[1] * O.party_encode(U_v, q, P), E_l + wts[2] * O.party_encode(U_l, q, P)]
    e = []
    for x, n in zip(em, "avl"):
        xr = O.ragged_pack(x, lengths)
        e.append(torch.cat([O.linear(xr, P[f"graph_net_{n}.weight"], P[f"graph_net_{n}.bias"]), xr], -1))
    if att == "concat_only":
        feat = torch.cat(e, -1)
    elif att == "lmf_only":
        feat = O.lmf_forward(e[0], e[1], e[2], P, prefix="lmf.")
    elif att == "gated":
        feat = O.mm_gated_attention(e[0], e[1], e[2], P)
    else:
        x = torch.zeros(T, len(lengths), 900)
        cat = torch.cat([e[2], e[0], e[1]], -1)
        off = 0
        for b, L in enumerate(lengths):
            x[:L, b] = cat[off:off + L]
            off += L
        out = O.mfn_forward(x, P, "mfn")
        feat = torch.cat([out[:L, b] for b, L in enumerate(lengths)], 0)
    lp_ref = torch.log_softmax(O.linear(feat * m_head.float() * 2.0, P["smax_fc.weight"], P["smax_fc.bias"]), 1)
    O.focal_loss(lp_ref, lab, 1.0).backward()
    # ---- kernels (no GRU masks injected = no inter-layer dropout; the fusion blocks' own Dropout layers in eval mode)
    m = m.to(DEV).train()
    for name in ("gatedatt", "mfn", "lmf"):
        if hasattr(m, name):
            getattr(m, name).eval()
    masks = {"head": m_head.to(DEV)}
    lp = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV), masks=masks)[0]
    assert float((lp.detach().cpu() - lp_ref.detach()).abs().max()) < 1e-4
    mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV)).backward()
    for k, p in m.named_parameters():
        if P[k].grad is None:
            assert p.grad is None, k
            continue
        g, r = p.grad.cpu(), P[k].grad
        assert float((g - r).norm() / max(float(r.norm()), 1e-8)) < 1e-3, k
    assert {n for n, _ in used_parameters(m)} == {k for k, p in m.named_parameters() if p.grad is not None}


def test_no_graph_tfn_only_runs_end_to_end():
    """att_type='tfn_only' (code/model.py:1389-1390): the model-level wiring of the TFN block (the block itself is pinned to
    the reference in tests/test_gpu_tfn.py): normalised log-probabilities, finite gradients for exactly the trainer's set."""
    import mmdfn_b200 as mm
    from mmdfn_b200.dp import used_parameters
    lengths, S, C = [7, 12, 5], 2, 6
    t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 40, 24, S, C, seed=41)
    torch.manual_seed(3)
    m = mm.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
                            window_future=10, n_classes=C, dropout=0.0, graph_type="None", D_m_v=24, D_m_a=40, modals="avl",
                            att_type="tfn_only", use_speaker=False, use_crn_speaker=True, speaker_weights="1-1-1").to(DEV).train()
    lp = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV))[0]
    assert lp.shape == (sum(lengths), C) and float((lp.exp().sum(1) - 1).abs().max()) < 1e-5
    mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV)).backward()
    got = {k for k, p in m.named_parameters() if p.grad is not None}
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    assert {n for n, _ in used_parameters(m)} == got


def test_graph_type_gf_is_gdf_without_the_fusion_gate():
    """graph_type='GF' (code/model.py:944-950): the same MM_GCN as 'GDF' built with reason_flag=False whatever the model's
    own flag says -- same logits as a 'GDF' model constructed with reason_flag=False and the same weights."""
    import mmdfn_b200 as mm
    lengths, S, C = [9, 14, 6], 2, 6
    t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 40, 24, S, C, seed=5)
    kw = dict(n_speakers=S, max_seq_len=200, window_past=10, window_future=10, n_classes=C, dropout=0.0, alpha=0.2, lamda=0.5,
              D_m_v=24, D_m_a=40, modals="avl", att_type="concat_subsequently", Deep_GCN_nlayers=2, use_speaker=False,
              use_crn_speaker=True, speaker_weights="3-0-1")
    m_gf = mm.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, graph_type="GF", reason_flag=True, **kw)
    m_gdf = mm.DialogueGNNModel("LSTM", 100, 150, 150, 100, 100, 100, 100, graph_type="GDF", reason_flag=False, **kw)
    assert not m_gf.graph_model.graph_net.reason_flag
    w = O.formula_weights({k: tuple(p.shape) for k, p in m_gf.state_dict().items()}, seed=7)
    m_gf.load_state_dict(w)
    m_gdf.load_state_dict(w)
    args = (t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV))
    lp1, lp2 = m_gf.to(DEV).eval()(*args)[0], m_gdf.to(DEV).eval()(*args)[0]
    assert torch.equal(lp1, lp2)
    with torch.no_grad():
        P = {k: p.detach().cpu() for k, p in m_gf.state_dict().items()}
        ref = O.forward_gdf(P, t, q, lengths, a, v, nlayers=2, speaker_weights=(3.0, 0.0, 1.0), reason_flag=False)
    assert float((lp1.detach().cpu() - ref).abs().max()) < 1e-4
