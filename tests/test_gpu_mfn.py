"""a13: the Memory Fusion Network block (code/model_fusion.py:10-120) on the GPU path -- mmdfn_mfn_fwd / _bwd through the
MFN module -- against (1) the unmodified reference module's output, input gradient and parameter-gradient summaries
stored in tests/golden/mfn.npz (eval mode) and (2) the oracle's restatement with autograd gradients, with and without
injected dropout masks; and DialogueGNNModel(graph_type='GDF', att_type='mfn') (code/model.py:1303-1330) against the
oracle.  Tolerances: 1e-5 on outputs, 2e-4 relative on gradients (fp32 sigmoid / tanh / softmax chains over T steps)."""
import numpy as np
import pytest
import torch

import mmdfn_oracle as O
from helpers import case_inputs, case_weights, load_case, spk_weights

pytestmark = pytest.mark.gpu
DEV = "cuda"


def maxerr(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def relerr(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _shapes():
    s = load_case("mfn")
    keys = [str(k) for k in s["keys"]]
    return s, keys, {k: tuple(int(v) for v in str(sh).strip("()").replace(" ", "").split(",") if v) for k, sh in zip(keys, s["shapes"])}


def _module(W):
    import mmdfn_b200
    m = mmdfn_b200.MFN()
    m.load_state_dict(W, strict=True)
    return m.to(DEV)


def test_state_dict_keys_match_the_reference_module():
    import mmdfn_b200
    s, keys, shapes = _shapes()
    sd = mmdfn_b200.MFN().state_dict()
    assert sorted(sd.keys()) == sorted(keys)           # (the golden file lists the keys sorted)
    assert list(sd.keys())[:4] == ['lstm_l.weight_ih', 'lstm_l.weight_hh', 'lstm_l.bias_ih', 'lstm_l.bias_hh'] and list(sd.keys())[-1] == 'out_fc2.bias'
    assert {k: tuple(v.shape) for k, v in sd.items()} == shapes


def test_matches_reference_golden():
    s, keys, shapes = _shapes()
    W = O.formula_weights(shapes, seed=5)
    m = _module(W).eval()
    x = torch.from_numpy(s["x"]).to(DEV).requires_grad_(True)
    out = m(x)
    assert out.shape == (9, 4, 400)
    assert float((out.detach().cpu() - torch.from_numpy(s["out"])).abs().max()) < 1e-5
    (out * torch.from_numpy(s["G"]).to(DEV)).sum().backward()
    assert float((x.grad.cpu() - torch.from_numpy(s["dx"])).abs().max()) < 2e-5
    P = dict(m.named_parameters())
    for k in keys:
        if not bool(s["used." + k]):
            assert P[k].grad is None, k                       # out_fc1 / out_fc2 never receive a gradient
            continue
        g = P[k].grad.cpu()
        assert abs(float(g.norm()) - float(s["gnorm." + k])) <= 3e-4 * max(1.0, float(s["gnorm." + k])), k
        assert abs(float(g.sum()) - float(s["gsum." + k])) <= 3e-4 * max(1.0, float(s["gnorm." + k])), k


@pytest.mark.parametrize("T,n,with_masks", [(1, 1, False), (9, 4, True), (23, 5, False), (40, 33, True)])
def test_forward_and_gradients_match_oracle(T, n, with_masks):
    s, keys, shapes = _shapes()
    W = O.formula_weights(shapes, seed=7)
    rs = np.random.RandomState(T * 100 + n)
    x = torch.from_numpy(rs.standard_normal((T, n, 900)).astype(np.float32))
    gout = torch.from_numpy(rs.standard_normal((T, n, 400)).astype(np.float32))
    keep = [rs.rand(T * n, 100) > 0.2 for _ in range(4)] if with_masks else None
    P = {k: w.clone().requires_grad_(True) for k, w in W.items()}
    xr = x.clone().requires_grad_(True)
    om = [torch.from_numpy(k.astype(np.float32) / 0.8).view(T, n, 100) for k in keep] if with_masks else None
    ref = O.mfn_forward(xr, P, masks=om)
    ref.backward(gout)
    m = _module(W).train() if with_masks else _module(W).eval()
    xd = x.clone().to(DEV).requires_grad_(True)
    out = m(xd, masks=[torch.from_numpy(k.astype(np.uint8)).to(DEV) for k in keep] if with_masks else None)
    out.backward(gout.to(DEV))
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 1e-5
    assert relerr(xd.grad, xr.grad) < 2e-4
    for k, prm in m.named_parameters():
        if P[k].grad is None:
            assert prm.grad is None, k
            continue
        assert relerr(prm.grad, P[k].grad) < 2e-4, k


def test_train_mode_draws_masks():
    s, keys, shapes = _shapes()
    m = _module(O.formula_weights(shapes, seed=7)).train()
    x = torch.randn(12, 3, 900, device=DEV)
    o1, o2 = m(x), m(x)
    assert torch.isfinite(o1).all() and not torch.equal(o1, o2)            # fresh Dropout(0.2) masks per call
    assert torch.equal(o1[..., :300], o2[..., :300])                       # the LSTM branches carry no dropout


def _mfn_model(c):
    import mmdfn_b200 as mm
    d = [int(x) for x in c["dims"]]
    S, C, K = int(c["S"]), int(c["C"]), int(c["K"])
    torch.manual_seed(11)
    m = mm.DialogueGNNModel(
        "LSTM", d[0], 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10, window_future=10,
        n_classes=C, dropout=0.4, nodal_attention=True, no_cuda=False, graph_type="GDF", alpha=0.2, lamda=0.5,
        multiheads=6, graph_construct="direct", use_GCN=False, use_residue=True, D_m_v=d[2], D_m_a=d[1], modals="avl",
        att_type="mfn", av_using_lstm=False, Deep_GCN_nlayers=K, dataset="IEMOCAP", use_speaker=False,
        use_modal=False, reason_flag=True, multi_modal=True, use_crn_speaker=True, speaker_weights=str(c["spk_w"]),
        modal_weight=1.0)
    W = case_weights(c)
    sd = m.state_dict()
    for k in sd:                       # the golden case's weights everywhere they exist; the MFN block and the 400-wide classifier keep their seeded init
        if k in W and tuple(W[k].shape) == tuple(sd[k].shape):
            sd[k] = W[k]
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("train", [False, True])
def test_model_with_mfn_head_against_oracle(train):
    """DialogueGNNModel(graph_type='GDF', att_type='mfn'): log-probabilities and every gradient against the oracle
    (the graph stack is the reference-pinned restatement, the head is oracle.mfn_head)."""
    import mmdfn_b200 as mm
    name = "c4_synth_small"
    c = load_case(name)
    t, a, v, q, u, lab, lengths = case_inputs(c, name)
    K, N, T, B = int(c["K"]), sum(lengths), t.shape[0], t.shape[1]
    m = _mfn_model(c)
    P = {k: w.detach().cpu().clone().requires_grad_(True) for k, w in m.state_dict().items()}
    om = gm = None
    if train:
        rs = np.random.RandomState(5)
        keep = [rs.rand(T * B, 100) > 0.2 for _ in range(4)]
        k_hd = rs.rand(N, 400) > 0.4
        om = {"mfn": [torch.from_numpy(k.astype(np.float32) / 0.8).view(T, B, 100) for k in keep],
              "mfn_head": torch.from_numpy(k_hd.astype(np.float32) / 0.6)}
        gm = {"mfn": [torch.from_numpy(k.astype(np.uint8)).to(DEV) for k in keep], "mfn_head": torch.from_numpy(k_hd.astype(np.uint8)).to(DEV)}
    lp_ref = O.forward_gdf(P, t, q, lengths, a, v, nlayers=K, speaker_weights=spk_weights(c), masks=om, att_type="mfn")
    loss_ref = O.focal_loss(lp_ref, lab, 1.0)
    loss_ref.backward()
    m = m.train() if train else m.eval()
    m.graph_model.graph_net.dropout = 0.0          # only the MFN block's and the head's dropout masks are injected
    lp = m(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV), masks=gm)[0]
    assert lp.shape == lp_ref.shape
    assert maxerr(lp, lp_ref) < 1e-4
    loss = mm.FocalLoss(gamma=1.0)(lp, lab.to(DEV))
    loss.backward()
    for k, pr in m.named_parameters():
        if P[k].grad is None:
            assert pr.grad is None, k
            continue
        assert relerr(pr.grad, P[k].grad) < 1e-3, k


def test_trainer_updates_the_mfn_parameters():
    """FlatAdamTrainer buckets exactly the parameters that receive a gradient (out_fc1 / out_fc2 stay outside)."""
    import mmdfn_b200 as mm
    from mmdfn_b200.dp import FlatAdamTrainer
    name = "c4_synth_small"
    c = load_case(name)
    t, a, v, q, u, lab, lengths = case_inputs(c, name)
    m = _mfn_model(c).train()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    tr = FlatAdamTrainer(m, mm.FocalLoss(gamma=1.0), lr=1e-3, weight_decay=0.0)
    loss = tr.step(t.to(DEV), q.to(DEV), u.to(DEV), lengths, a.to(DEV), v.to(DEV), lab.to(DEV))
    assert bool(torch.isfinite(loss))
    moved = {k for k, p in m.named_parameters() if not torch.equal(p.detach(), before[k])}
    assert "mfn.lstm_l.weight_hh" in moved and "mfn.gamma2_fc2.bias" in moved and "smax_fc.weight" in moved
    assert not any(k.startswith("mfn.out_fc") for k in moved)
