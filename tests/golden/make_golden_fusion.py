"""Golden fixtures for the remaining fusion blocks of code/model_fusion.py -- LMF (:214-310) and TFN (:123-211) -- from the
UNMODIFIED reference modules on CPU fp32.  Run in the build container only:
    python tests/golden/make_golden_fusion.py        -> tests/golden/lmf.npz, tests/golden/tfn.npz
Weights = oracle.formula_weights of the state_dict shapes (LMF seed 13, TFN seed 17: the 309 M-element post-fusion weight
of TFN is regenerated from the formula by the tests, not stored); eval mode (Dropout = identity); stored: the output, the
three input gradients and per-parameter gradient summaries under a fixed cotangent."""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
O = MG.O


def run(mod, seed, n, path, big=()):
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    mod.load_state_dict(O.formula_weights(shapes, seed=seed), strict=True)
    mod.eval()
    rs = np.random.RandomState(seed)
    xs = [torch.from_numpy((0.5 * rs.standard_normal((n, 300))).astype(np.float32)).requires_grad_(True) for _ in range(3)]
    out = mod(*xs)
    G = torch.from_numpy(rs.standard_normal(tuple(out.shape)).astype(np.float32))
    (out * G).sum().backward()
    fix = {"xa": xs[0].detach().numpy(), "xv": xs[1].detach().numpy(), "xt": xs[2].detach().numpy(), "G": G.numpy(),
           "out": out.detach().numpy(), "dxa": xs[0].grad.numpy(), "dxv": xs[1].grad.numpy(), "dxt": xs[2].grad.numpy()}
    for k, p in mod.named_parameters():
        fix["gnorm." + k] = np.array(float(p.grad.norm()))
        fix["gsum." + k] = np.array(float(p.grad.double().sum()))
        if k in big:                                   # a fixed 4096-element sample of a huge gradient instead of nothing
            idx = np.random.RandomState(1).randint(0, p.numel(), size=4096)
            fix["gidx." + k] = idx
            fix["gval." + k] = p.grad.reshape(-1)[torch.from_numpy(idx)].numpy()
    fix["keys"] = np.array(sorted(shapes))
    fix["shapes"] = np.array([str(shapes[k]) for k in sorted(shapes)])
    np.savez_compressed(path, **fix)
    print(os.path.basename(path), tuple(out.shape), float(out.abs().mean()))


def main():
    MG.install_shim()
    import model_fusion
    torch.manual_seed(0)
    run(model_fusion.LMF(), 13, 11, os.path.join(HERE, "lmf.npz"))
    if "--tfn" in sys.argv:
        run(model_fusion.TFN(), 17, 5, os.path.join(HERE, "tfn.npz"), big=("post_fusion_layer_1.weight",))


if __name__ == "__main__":
    main()
