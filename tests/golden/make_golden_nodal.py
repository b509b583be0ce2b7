"""Golden fixture for the nodal-attention head of the `relation` graph type:
`classify_node_features(..., nodal_attn=True)` -> `attentive_node_features` -> `MatchingAttention('general2')`
(code/model.py:614-672, 31-86), from the UNMODIFIED reference functions on CPU fp32.  Run in the build container only:
    python tests/golden/make_golden_nodal.py        -> tests/golden/nodal_head.npz
Inputs: ragged node features (N, 300) of 5 dialogues (lengths incl. 1 and a longest one that defines the padding),
umask (B, T); layers MatchingAttention(300, 300, 'general2'), Linear(300, 100), Dropout(0) (identity, so that the
train-mode backward is the eval-mode function), Linear(100, 6); weights = oracle.formula_weights (seed 9).  Stored: the
log-probabilities, the gradient w.r.t. the features and per-parameter gradient summaries under a fixed cotangent."""
import os, sys
import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
O = MG.O

LENGTHS = [7, 1, 12, 5, 9]
D, HID, C = 300, 100, 6


def main():
    MG.install_shim()
    import model as RM
    torch.manual_seed(0)
    layers = {"matchatt": RM.MatchingAttention(D, D, att_type='general2'), "linear": nn.Linear(D, HID), "smax_fc": nn.Linear(HID, C)}
    shapes = {f"{n}.{k}": tuple(v.shape) for n, m in layers.items() for k, v in m.state_dict().items()}
    w = O.formula_weights(shapes, seed=9)
    for n, m in layers.items():
        m.load_state_dict({k: w[f"{n}.{k}"] for k in m.state_dict()}, strict=True)
    rs = np.random.RandomState(23)
    N, T, B = sum(LENGTHS), max(LENGTHS), len(LENGTHS)
    x = torch.from_numpy((0.5 * rs.standard_normal((N, D))).astype(np.float32)).requires_grad_(True)
    umask = torch.zeros(B, T)
    for b, L in enumerate(LENGTHS):
        umask[b, :L] = 1
    G = torch.from_numpy(rs.standard_normal((N, C)).astype(np.float32))
    lp = RM.classify_node_features(x, LENGTHS, umask, layers["matchatt"], layers["linear"], nn.Dropout(0.0), layers["smax_fc"],
                                   True, False, True)
    (lp * G).sum().backward()
    att = RM.attentive_node_features(x.detach(), LENGTHS, umask, layers["matchatt"], True)      # (T, B, D), padded rows included
    fix = {"lengths": np.array(LENGTHS), "x": x.detach().numpy(), "G": G.numpy(), "log_prob": lp.detach().numpy(),
           "dx": x.grad.numpy(), "att_padded": att.detach().numpy()}
    for n, m in layers.items():
        for k, p in m.named_parameters():
            fix[f"w.{n}.{k}"] = p.detach().numpy()
            fix[f"gnorm.{n}.{k}"] = np.array(float(p.grad.norm()))
            fix[f"gsum.{n}.{k}"] = np.array(float(p.grad.sum()))
    np.savez_compressed(os.path.join(HERE, "nodal_head.npz"), **fix)
    print("nodal head golden:", tuple(lp.shape), float(lp.mean()), float(x.grad.abs().mean()))


if __name__ == "__main__":
    main()
