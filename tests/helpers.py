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


def case_inputs(c, name):
    """(textf, acouf, visuf, qmask, umask, label, lengths) tensors of a golden case."""
    lengths = [int(x) for x in c["lengths"]]
    if "textf" in c:
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
