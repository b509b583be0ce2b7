"""Golden fixture for the ablation fusion block MFN (code/model_fusion.py:10-120, `att_type='mfn'`): output and gradient
summaries of the UNMODIFIED reference module on CPU fp32 (its hard-coded .cuda() calls are mapped to the identity by the
shim of make_golden.py).  Run in the build container only:
    python tests/golden/make_golden_mfn.py        -> tests/golden/mfn.npz
Weights are `oracle.formula_weights` of the state_dict shapes (seed 5); eval mode (the module's five Dropout(0.2) are
identities), so gradients come from an eval-mode backward (no in-place ops in this module)."""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (shim + oracle import path)
O = MG.O


def main():
    MG.install_shim()
    import model_fusion
    torch.manual_seed(0)
    mfn = model_fusion.MFN()
    shapes = {k: tuple(v.shape) for k, v in mfn.state_dict().items()}
    mfn.load_state_dict(O.formula_weights(shapes, seed=5), strict=True)
    mfn.eval()
    rs = np.random.RandomState(17)
    T, n = 9, 4
    x = torch.from_numpy(rs.standard_normal((T, n, 900)).astype(np.float32)).requires_grad_(True)
    G = torch.from_numpy(rs.standard_normal((T, n, 400)).astype(np.float32))
    out = mfn(x)
    (out * G).sum().backward()
    fix = {"x": x.detach().numpy(), "G": G.numpy(), "out": out.detach().numpy(), "dx": x.grad.numpy()}
    for k, p in mfn.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)      # out_fc1/out_fc2 are never used by forward()
        fix["gnorm." + k] = np.array(float(g.norm()))
        fix["gsum." + k] = np.array(float(g.sum()))
        fix["used." + k] = np.array(p.grad is not None)
    fix["keys"] = np.array(sorted(shapes))
    fix["shapes"] = np.array([str(shapes[k]) for k in sorted(shapes)])
    np.savez_compressed(os.path.join(HERE, "mfn.npz"), **fix)
    print("mfn golden:", out.shape, float(out.abs().mean()))


if __name__ == "__main__":
    main()
