"""f4: the LMF fusion block (code/model_fusion.py:214-310) on the GPU path against (1) the unmodified reference module
(tests/golden/lmf.npz: output, the three input gradients, parameter-gradient summaries) and (2) the oracle's autograd on a
larger batch.  Tolerances: 1e-5 on outputs, 2e-4 relative on gradients."""
import os

import numpy as np
import pytest
import torch

import mmdfn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def _module(seed):
    from mmdfn_b200.modules import LMF
    m = LMF()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    P = O.formula_weights(shapes, seed=seed)
    m.load_state_dict(P, strict=True)
    return m.to(DEV), P


def test_state_dict_and_init_match_reference_layout():
    from mmdfn_b200.modules import LMF
    g = np.load(os.path.join(HERE, "golden", "lmf.npz"))
    m = LMF()
    assert sorted(m.state_dict().keys()) == list(g["keys"])
    assert [str(tuple(m.state_dict()[k].shape)) for k in sorted(m.state_dict())] == list(g["shapes"])


def test_matches_reference_golden():
    g = np.load(os.path.join(HERE, "golden", "lmf.npz"))
    m, _ = _module(13)
    xs = [torch.from_numpy(g[k]).to(DEV).requires_grad_(True) for k in ("xa", "xv", "xt")]
    out = m(*xs)
    assert float((out.detach().cpu() - torch.from_numpy(g["out"])).abs().max()) < 1e-5
    (out * torch.from_numpy(g["G"]).to(DEV)).sum().backward()
    for x, k in zip(xs, ("dxa", "dxv", "dxt")):
        assert float((x.grad.cpu() - torch.from_numpy(g[k])).abs().max()) < 2e-5 * max(1.0, float(np.abs(g[k]).max()))
    for k, p in m.named_parameters():
        ref = float(g["gnorm." + k])
        assert abs(float(p.grad.norm()) - ref) < 2e-4 * max(1.0, ref), k
        assert abs(float(p.grad.double().sum()) - float(g["gsum." + k])) < 2e-4 * max(1.0, ref), k


@pytest.mark.parametrize("N", [1, 257, 1623])
def test_forward_and_gradients_match_oracle(N):
    m, P = _module(29)
    rs = np.random.RandomState(N)
    xs0 = [torch.from_numpy((0.5 * rs.standard_normal((N, 300))).astype(np.float32)) for _ in range(3)]
    G = torch.from_numpy(rs.standard_normal((N, 300)).astype(np.float32))
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    xr = [x.clone().requires_grad_(True) for x in xs0]
    ref = O.lmf_forward(*xr, Pr)
    (ref * G).sum().backward()
    xs = [x.to(DEV).requires_grad_(True) for x in xs0]
    out = m(*xs)
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
    (out * G.to(DEV)).sum().backward()
    for x, r in zip(xs, xr):
        assert float((x.grad.cpu() - r.grad).norm() / max(float(r.grad.norm()), 1e-12)) < 2e-4
    for k, p in m.named_parameters():
        r = Pr[k].grad
        assert float((p.grad.cpu() - r).norm() / max(float(r.norm()), 1e-12)) < 2e-4, k
