"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/code, imported through the shim of SURVEY.md Appendix A) on CPU fp32.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Weights are not stored: they are `oracle.formula_weights` of the state_dict shapes
(deterministic, numpy-only), loaded into the reference with load_state_dict(strict=True).
Synthetic inputs are `oracle.synthetic_batch` (seeded); only real-data inputs are stored.
"""
import os, sys, types, pickle
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmdfn_oracle as O  # noqa: E402


def install_shim():
    sys.path.insert(0, os.path.join(REF, "code"))
    tg, tgnn = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn")

    class _NA(torch.nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError("torch_geometric not installed")

    tgnn.RGCNConv = tgnn.GraphConv = _NA
    tg.nn = tgnn
    sys.modules["torch_geometric"], sys.modules["torch_geometric.nn"] = tg, tgnn
    torch.Tensor.cuda = lambda self, *a, **k: self
    _si, _gi = torch.Tensor.__setitem__, torch.Tensor.__getitem__
    fix = lambda i: tuple(torch.as_tensor(r) for r in i) if isinstance(i, np.ndarray) and i.ndim == 2 else i
    torch.Tensor.__setitem__ = lambda self, i, v: _si(self, fix(i), v)
    torch.Tensor.__getitem__ = lambda self, i: _gi(self, fix(i))


def identity_dropout(model):
    """train()-mode gradients with dropout = identity (eval-mode backward raises)."""
    import torch.nn.functional as F
    F.dropout = lambda x, p=0.5, training=True, inplace=False: x.clone()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.GRU):
            m.dropout = 0.0
    model.train()


def make_model(model_mod, d_text, d_audio, d_visual, S, C, K, dataset, spk_w, reason_flag=True, graph_type="GDF"):
    m = model_mod.DialogueGNNModel(
        "LSTM", d_text, 150, 150, 100, 100, 100, 100, n_speakers=S, max_seq_len=200, window_past=10,
        window_future=10, n_classes=C, dropout=0.4, nodal_attention=True, no_cuda=True, graph_type=graph_type,
        alpha=0.2, lamda=0.5, multiheads=6, graph_construct="direct", use_GCN=False, use_residue=True,
        D_m_v=d_visual, D_m_a=d_audio, modals="avl", att_type="concat_subsequently", av_using_lstm=False,
        Deep_GCN_nlayers=K, dataset=dataset, use_speaker=False, use_modal=False, reason_flag=reason_flag,
        multi_modal=True, use_crn_speaker=True, speaker_weights=spk_w, modal_weight=1.0)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(O.formula_weights(shapes), strict=True)
    return m, shapes


def grad_summary(model):
    out = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().reshape(-1).double()
        rs = np.random.RandomState(len(name) * 7919 + g.numel())
        proj = torch.from_numpy(rs.standard_normal(g.numel()))
        out[name] = np.array([float(g.norm()), float(g.sum()), float((g * proj).sum())], dtype=np.float64)
    return out


def collate(samples):
    from torch.nn.utils.rnn import pad_sequence
    cols = list(zip(*samples))
    return [pad_sequence(list(cols[i])) if i < 4 else pad_sequence(list(cols[i]), True) for i in range(6)]


def run_case(name, model_mod, loss_mod, batch, cfg, train_grads, class_weights=None, gamma=1.0, store_inputs=True):
    textf, visuf, acouf, qmask, umask, label = batch
    lengths = [int(umask[j].sum().item()) for j in range(umask.shape[0])]
    m, shapes = make_model(model_mod, textf.shape[2], acouf.shape[2], visuf.shape[2], cfg["S"], cfg["C"],
                           cfg["K"], cfg["dataset"], cfg["spk_w"])
    out = {"lengths": np.array(lengths, np.int64), "K": cfg["K"], "S": cfg["S"], "C": cfg["C"],
           "spk_w": cfg["spk_w"], "dims": np.array([textf.shape[2], acouf.shape[2], visuf.shape[2]])}
    if store_inputs:
        out.update(textf=textf.numpy(), acouf=acouf.numpy(), visuf=visuf.numpy(), qmask=qmask.numpy(),
                   umask=umask.numpy())
    lab = torch.cat([label[j][:lengths[j]] for j in range(len(lengths))])
    out["label"] = lab.numpy()
    m.eval()
    with torch.no_grad():
        lp = m(textf, qmask, umask, lengths, acouf, visuf)[0]
    out["log_prob_eval"] = lp.numpy()
    if train_grads:
        identity_dropout(m)
        lp = m(textf, qmask, umask, lengths, acouf, visuf)[0]
        loss = loss_mod.FocalLoss(gamma=gamma, alpha=class_weights)(lp, lab)
        loss.backward()
        out["log_prob_train"] = lp.detach().numpy()
        out["loss"] = np.array(float(loss))
        out["gamma"] = np.array(gamma)
        if class_weights is not None:
            out["class_weights"] = class_weights.numpy()
        for k, v in grad_summary(m).items():
            out["grad::" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N =", sum(lengths), "lengths", lengths, "log_prob[0]", out["log_prob_eval"][0])


def main():
    install_shim()
    import model as model_mod, loss as loss_mod, model_mm, model_GCN  # reference modules
    import torch.nn.functional as F
    real_dropout = F.dropout

    cw_ie = torch.FloatTensor([1 / 0.086747, 1 / 0.144406, 1 / 0.227883, 1 / 0.160585, 1 / 0.127711, 1 / 0.252668])

    # ---- real data -------------------------------------------------------------------------
    ie = pickle.load(open(os.path.join(REF, "data/iemocap/IEMOCAP_features.pkl"), "rb"), encoding="latin1")
    ids, spk, labels, text, audio, visual, sent, train_vid, test_vid = ie

    def ie_sample(vid):
        return (torch.FloatTensor(text[vid]), torch.FloatTensor(visual[vid]), torch.FloatTensor(audio[vid]),
                torch.FloatTensor([[1, 0] if x == "M" else [0, 1] for x in spk[vid]]),
                torch.FloatTensor([1] * len(labels[vid])), torch.LongTensor(labels[vid]))

    cfg_ie = dict(S=2, C=6, dataset="IEMOCAP", spk_w="3-0-1")
    first_train = [x for x in train_vid][0]
    run_case("c1_iemocap_single", model_mod, loss_mod, collate([ie_sample(first_train)]), dict(cfg_ie, K=1), False)
    short = sorted([x for x in test_vid], key=lambda v: (len(labels[v]), v))[:4]
    run_case("c2_iemocap_b4", model_mod, loss_mod, collate([ie_sample(v) for v in short]), dict(cfg_ie, K=2), True,
             class_weights=cw_ie, gamma=1.0)
    F.dropout = real_dropout

    me = pickle.load(open(os.path.join(REF, "data/meld/MELD_features_raw1.pkl"), "rb"), encoding="latin1")
    ids, spk, labels, text, audio, visual, sent, train_vid, test_vid, _ = me

    def me_sample(vid):
        return (torch.FloatTensor(text[vid]), torch.FloatTensor(visual[vid]), torch.FloatTensor(audio[vid]),
                torch.FloatTensor(spk[vid]), torch.FloatTensor([1] * len(labels[vid])), torch.LongTensor(labels[vid]))

    keys = sorted([x for x in test_vid])
    # 8 dialogues incl. a length-1 one and ones with many distinct speakers
    by_nspk = sorted(keys, key=lambda v: (-len({tuple(s) for s in spk[v]}), len(labels[v]), v))
    pick = by_nspk[:3] + [k for k in keys if len(labels[k]) == 1][:1] + [k for k in keys if 5 <= len(labels[k]) <= 9][:4]
    run_case("c3_meld_b8", model_mod, loss_mod, collate([me_sample(v) for v in pick]),
             dict(S=9, C=7, K=4, dataset="MELD", spk_w="0.5-0.5-1.5"), True, class_weights=None, gamma=1.0)
    F.dropout = real_dropout

    # ---- synthetic (inputs regenerated from the seed, not stored) ------------------------------
    def synth(lengths, S, C, seed):
        t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 512, 1024, S, C, seed)
        pos = np.cumsum([0] + list(lengths))
        lab_pad = torch.zeros(len(lengths), max(lengths), dtype=torch.long)
        for b, L in enumerate(lengths):
            lab_pad[b, :L] = lab[pos[b]:pos[b + 1]]
        return [t, v, a, q, u, lab_pad]

    run_case("c4_synth_small", model_mod, loss_mod, synth([9, 14, 6], 2, 6, 4),
             dict(S=2, C=6, K=2, dataset="IEMOCAP", spk_w="3-0-1"), True, class_weights=cw_ie, gamma=0.5,
             store_inputs=False)
    F.dropout = real_dropout
    run_case("c5_synth_small", model_mod, loss_mod, synth([40, 25], 8, 6, 5),
             dict(S=8, C=6, K=6, dataset="IEMOCAP", spk_w="1-1-1"), True, class_weights=None, gamma=1.0,
             store_inputs=False)
    F.dropout = real_dropout

    # ---- sub-module goldens ---------------------------------------------------------------------
    sub = {}
    rs = np.random.RandomState(11)
    dia = [5, 3, 7]
    N = sum(dia)
    a, v, l = (torch.from_numpy(rs.standard_normal((N, 200)).astype(np.float32)) for _ in range(3))
    mm = model_mm.MM_GCN(200, 200, 200, 200, 3, 100, 6, 0.4, 0.5, 0.2, True, True, True, n_speakers=2,
                         modals=["a", "v", "l"], use_speaker=False, use_modal=False, reason_flag=True, modal_weight=1.0)
    shp = {k: tuple(t.shape) for k, t in mm.state_dict().items()}
    mm.load_state_dict(O.formula_weights(shp, seed=5))
    mm.eval()
    with torch.no_grad():
        adj = mm.create_big_adj(a, v, l, dia, ["a", "v", "l"], 1.0)
        qm = torch.zeros(max(dia), len(dia), 2); qm[:, :, 0] = 1
        feat = mm(a.clone(), v.clone(), l.clone(), dia, qm)
        x = torch.cat([a, v, l], 0)
        gout = mm.graph_net(x, None, None, adj)
        h0 = torch.relu(mm.graph_net.fcs[0](x))
        conv = mm.graph_net.convs[1](h0, adj, h0, 0.5, 0.2, 2)
    sub.update(adj_a=a.numpy(), adj_v=v.numpy(), adj_l=l.numpy(), adj_dia=np.array(dia), adj_dense=adj.numpy(),
               mmgcn_out=feat.numpy(), gcnii_out=gout.numpy(), conv_in=h0.numpy(), conv_out=conv.numpy())
    # adjacency with modal_weight != 1
    with torch.no_grad():
        sub["adj_dense_mw"] = mm.create_big_adj(a, v, l, dia, ["a", "v", "l"], 0.7).numpy()
    # focal loss
    lp = torch.log_softmax(torch.from_numpy(rs.standard_normal((17, 6)).astype(np.float32)), 1)
    tg = torch.from_numpy(rs.randint(0, 6, size=17).astype(np.int64))
    sub.update(fl_lp=lp.numpy(), fl_tg=tg.numpy(),
               fl_g0=np.array(float(loss_mod.FocalLoss(gamma=0)(lp, tg))),
               fl_g1w=np.array(float(loss_mod.FocalLoss(gamma=1, alpha=cw_ie)(lp, tg))),
               fl_g05sum=np.array(float(loss_mod.FocalLoss(gamma=0.5, alpha=cw_ie, size_average=False)(lp, tg))))
    # relation-path: batch_graphify (edges/types/norms) with MaskedEdgeAttention
    for tag, S, lens in (("rel2", 2, [37, 12, 50]), ("rel9", 9, [5, 1, 23])):
        T, B = max(lens), len(lens)
        feats = torch.from_numpy(rs.standard_normal((T, B, 200)).astype(np.float32))
        spk_id = rs.randint(0, S, size=(T, B))
        qm = np.zeros((T, B, S), np.float32)
        for b, L in enumerate(lens):
            qm[np.arange(L), b, spk_id[:L, b]] = 1
        att = model_mod.MaskedEdgeAttention(200, 200, True)
        att.load_state_dict(O.formula_weights({k: tuple(t.shape) for k, t in att.state_dict().items()}, seed=9))
        etm = {}
        for j in range(S):
            for k in range(S):
                etm[str(j) + str(k) + "0"] = len(etm)
                etm[str(j) + str(k) + "1"] = len(etm)
        with torch.no_grad():
            nf, ei, en, et, eil = model_mod.batch_graphify(feats, torch.from_numpy(qm), lens, 10, 10, etm, att, True)
        ei, en, et = ei.numpy(), en.numpy(), et.numpy()
        order = np.lexsort((ei[1], ei[0]))
        sub.update({f"{tag}_feats": feats.numpy(), f"{tag}_qmask": qm, f"{tag}_lens": np.array(lens),
                    f"{tag}_att_w": att.scalar.weight.detach().numpy(), f"{tag}_edge_index": ei[:, order],
                    f"{tag}_edge_norm": en[order], f"{tag}_edge_type": et[order], f"{tag}_edge_lens": np.array(eil),
                    f"{tag}_node_features": nf.numpy()})
    # MMGatedAttention ('general'), eval mode
    ga = model_mod.MMGatedAttention(300, 100, att_type="general")
    gshp = {k: tuple(t.shape) for k, t in ga.state_dict().items()}
    ga.load_state_dict(O.formula_weights(gshp, seed=3))
    ga.eval()
    xa, xv, xl = (torch.from_numpy(rs.standard_normal((11, 300)).astype(np.float32)) for _ in range(3))
    with torch.no_grad():
        sub.update(ga_a=xa.numpy(), ga_v=xv.numpy(), ga_l=xl.numpy(), ga_out=ga(xa, xv, xl, ["a", "v", "l"]).numpy())
    np.savez_compressed(os.path.join(HERE, "submodules.npz"), **sub)
    # state_dict key/shape manifest of the reference for the drop-in boundary test
    manifest = {}
    for tag, args in (("iemocap_k2", (100, 1582, 342, 2, 6, 2, "IEMOCAP", "3-0-1")),
                      ("meld_k4", (600, 300, 342, 9, 7, 4, "MELD", "0.5-0.5-1.5"))):
        _, shapes = make_model(model_mod, *args)
        manifest[tag] = {k: list(v) for k, v in shapes.items()}
    import json
    json.dump(manifest, open(os.path.join(HERE, "state_dict_manifest.json"), "w"), indent=0)
    print("done")


if __name__ == "__main__":
    main()
