"""Shared helpers for the parity tests (test infrastructure; may use the oracle)."""
import json
import os

import numpy as np
import torch

import mmdfn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def manifest(tag):
    return {k: tuple(v) for k, v in json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))[tag].items()}


def model_shapes(d_text, d_audio, d_visual, S, C, K):
    """state_dict shapes of DialogueGNNModel (GDF / LSTM / avl) -- derived from the IEMOCAP
    manifest of the reference by substituting the size-dependent entries."""
    base = manifest("iemocap_k2")
    out = {}
    for k, shp in base.items():
        if k.startswith("graph_model.graph_net.convs."):
            continue
        out[k] = shp
    out["linear_l.weight"] = (200, d_text)
    out["linear_a.weight"] = (200, d_audio)
    out["linear_v.weight"] = (200, d_visual)
    for i in range(K):
        out[f"graph_model.graph_net.convs.{i}.weight"] = (200, 100)
    for k in ("graph_model.speaker_embeddings.weight", "graph_model.a_spk_embs.weight",
              "graph_model.v_spk_embs.weight", "graph_model.l_spk_embs.weight"):
        out[k] = (S, 200)
    out["graph_model.final_fc.weight"] = (C, 100)
    out["graph_model.final_fc.bias"] = (C,)
    out["smax_fc.weight"] = (C, 900)
    out["smax_fc.bias"] = (C,)
    return out


_STAGED = {}


def iemocap_batch(vids):
    """the reference's collate of these IEMOCAP test dialogues (code/dataloader.py:18-34): time-major zero-padded
    features, one-hot qmask ('M' -> [1,0]), umask.  Raw features come from baseline/_ref/data/iemocap_test_dialogues.npz
    (git-ignored; written by __graft_entry__.stage_reference() from the reference's pickle; travels to the GPU box) or
    straight from the pickle when /root/reference exists.  None when neither is there."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    npz = os.path.join(root, "baseline", "_ref", "data", "iemocap_test_dialogues.npz")
    pkl = "/root/reference/data/iemocap/IEMOCAP_features.pkl"
    if "src" not in _STAGED:
        if os.path.exists(npz):
            z = np.load(npz, allow_pickle=False)
            _STAGED["src"] = lambda kind, v: z[kind + "::" + v]
        elif os.path.exists(pkl):
            import pickle
            ids, spk, labels, text, audio, visual = pickle.load(open(pkl, "rb"), encoding="latin1")[:6]
            tab = {"text": text, "audio": audio, "visual": visual, "label": labels}
            _STAGED["src"] = lambda kind, v: (np.array([0 if x == "M" else 1 for x in spk[v]]) if kind == "spk"
                                              else np.asarray(tab[kind][v], np.float32 if kind != "label" else np.int64))
        else:
            _STAGED["src"] = None
    src = _STAGED["src"]
    if src is None:
        return None
    lengths = [len(src("label", v)) for v in vids]
    T, B = max(lengths), len(vids)

    def pad(kind):
        d = src(kind, vids[0]).shape[1]
        out = np.zeros((T, B, d), np.float32)
        for b, v in enumerate(vids):
            out[:lengths[b], b] = src(kind, v)
        return torch.from_numpy(out)

    q = np.zeros((T, B, 2), np.float32)
    u = np.zeros((B, T), np.float32)
    for b, v in enumerate(vids):
        q[np.arange(lengths[b]), b, src("spk", v)] = 1
        u[b, :lengths[b]] = 1
    return pad("text"), pad("audio"), pad("visual"), torch.from_numpy(q), torch.from_numpy(u)


def case_inputs(c, name):
    """(textf, acouf, visuf, qmask, umask, label, lengths) tensors of a golden case."""
    lengths = [int(x) for x in c["lengths"]]
    if "vids" in c:                      # real-size case: inputs come from the staged pickle (see make_golden_real.py)
        got = iemocap_batch([str(x) for x in c["vids"]])
        if got is None:
            import pytest
            pytest.skip("IEMOCAP test dialogues not staged under baseline/_ref/data (run __graft_entry__.build() where /root/reference exists)")
        t, a, v, q, u = got
    elif "textf" in c:
        t, a, v, q, u = (torch.from_numpy(c[k]) for k in ("textf", "acouf", "visuf", "qmask", "umask"))
    else:
        seed = {"c4_synth_small": 4, "c5_synth_small": 5}[name]
        t, a, v, q, u, _ = O.synthetic_batch(lengths, 100, 512, 1024, int(c["S"]), int(c["C"]), seed)
    return t, a, v, q, u, torch.from_numpy(c["label"]), lengths


def case_weights(c):
    d = [int(x) for x in c["dims"]]
    return O.formula_weights(model_shapes(d[0], d[1], d[2], int(c["S"]), int(c["C"]), int(c["K"])))


def grad_summary_of(named_grads):
    out = {}
    for name, g in named_grads.items():
        if g is None:
            continue
        g = g.detach().reshape(-1).double().cpu()
        rs = np.random.RandomState(len(name) * 7919 + g.numel())
        proj = torch.from_numpy(rs.standard_normal(g.numel()))
        out[name] = np.array([float(g.norm()), float(g.sum()), float((g * proj).sum())])
    return out


def spk_weights(c):
    return tuple(float(x) for x in str(c["spk_w"]).split("-"))


def write_small_iemocap_pickle(path, n_train=24):
    """A pickle in the author's IEMOCAP format (code/dataloader.py:12-14) holding the 31 staged IEMOCAP test dialogues:
    the first `n_train` form the train split, the rest the test split.  Returns `path` (None if nothing is staged)."""
    import pickle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    npz = os.path.join(root, "baseline", "_ref", "data", "iemocap_test_dialogues.npz")
    if not os.path.exists(npz):
        return None
    z = np.load(npz, allow_pickle=False)
    vids = [str(v) for v in z["vids"]]
    ids = {v: list(range(len(z["label::" + v]))) for v in vids}
    spk = {v: ["M" if s == 0 else "F" for s in z["spk::" + v]] for v in vids}
    labels = {v: [int(x) for x in z["label::" + v]] for v in vids}
    text = {v: [row for row in z["text::" + v]] for v in vids}
    audio = {v: [row for row in z["audio::" + v]] for v in vids}
    visual = {v: [row for row in z["visual::" + v]] for v in vids}
    sent = {v: [""] * len(labels[v]) for v in vids}
    pickle.dump((ids, spk, labels, text, audio, visual, sent, vids[:n_train], vids[n_train:]), open(path, "wb"))
    return path
