"""Golden fixture for GCNII (code/model_GCN.py:224-306), the per-modality deep GCN of graph_type='DeepGCN', from the
UNMODIFIED reference class on CPU fp32.  Run in the build container only:
    python tests/golden/make_golden_gcnii.py        -> tests/golden/gcnii.npz
Two cases on the same ragged input (N, 200): reason_flag=True, K=3, eval mode: output only (the reference's
`layer_inner += q` modifies the ReLU output in place, so its backward raises in every mode); reason_flag=False, K=4, train
mode with dropout 0: output, input gradient and parameter-gradient summaries.  lamda=0.5, alpha=0.1 (code/model.py:930-939)."""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
O = MG.O
LENGTHS = [9, 1, 17, 6]


def main():
    MG.install_shim()
    import model_GCN
    rs = np.random.RandomState(29)
    N = sum(LENGTHS)
    x0 = (0.7 * rs.standard_normal((N, 200))).astype(np.float32)
    G = rs.standard_normal((N, 300)).astype(np.float32)
    fix = {"lengths": np.array(LENGTHS), "x": x0, "G": G}
    for tag, K, reason in (("r", 3, True), ("n", 4, False)):
        torch.manual_seed(0)
        net = model_GCN.GCNII(nfeat=200, nlayers=K, nhidden=100, nclass=6, dropout=0.0, lamda=0.5, alpha=0.1, variant=True,
                              return_feature=True, use_residue=True, reason_flag=reason)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(O.formula_weights(shapes, seed=61 + K), strict=True)
        x = torch.from_numpy(x0).requires_grad_(True)
        if reason:
            net.eval()
            with torch.no_grad():
                out = net(x, LENGTHS, None)
        else:
            net.train()
            out = net(x, LENGTHS, None)
            (out * torch.from_numpy(G)).sum().backward()
            fix[tag + ".dx"] = x.grad.numpy()
            for k, p in net.named_parameters():
                fix[f"{tag}.used.{k}"] = np.array(p.grad is not None)
                if p.grad is not None:
                    fix[f"{tag}.gnorm.{k}"] = np.array(float(p.grad.norm()))
                    fix[f"{tag}.gsum.{k}"] = np.array(float(p.grad.sum()))
        fix[tag + ".out"] = out.detach().numpy()
        fix[tag + ".keys"] = np.array(sorted(shapes))
        fix[tag + ".shapes"] = np.array([str(shapes[k]) for k in sorted(shapes)])
        print(tag, tuple(out.shape), float(out.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "gcnii.npz"), **fix)


if __name__ == "__main__":
    main()
