"""a13: MMGatedAttention ('general') on the GPU path against (1) the output of the unmodified reference module stored in
tests/golden (eval mode) and (2) the oracle's restatement with autograd gradients, with and without injected dropout
masks.  Tolerances: 1e-5 on outputs (fp32 tanh / sigmoid / GEMM), 2e-4 relative on gradients."""
import numpy as np
import pytest
import torch

import mmdfn_oracle as O
from helpers import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = {"gatedatt.transform_l.weight": (100, 300), "gatedatt.transform_l.bias": (100,),
          "gatedatt.transform_v.weight": (100, 300), "gatedatt.transform_v.bias": (100,),
          "gatedatt.transform_a.weight": (100, 300), "gatedatt.transform_a.bias": (100,),
          "gatedatt.transform_av.weight": (1, 900), "gatedatt.transform_av.bias": (1,),
          "gatedatt.transform_al.weight": (1, 900), "gatedatt.transform_al.bias": (1,),
          "gatedatt.transform_vl.weight": (1, 900), "gatedatt.transform_vl.bias": (1,)}


def _module(P):
    import mmdfn_b200
    m = mmdfn_b200.MMGatedAttention(300, 100, att_type='general')
    m.load_state_dict({k.split(".", 1)[1]: v for k, v in P.items()}, strict=True)
    return m.to(DEV)


def test_matches_reference_golden_in_eval_mode():
    s = load_case("submodules")
    P = O.formula_weights(SHAPES, seed=3)
    m = _module(P).eval()
    a, v, l = (torch.from_numpy(s[k]).to(DEV) for k in ("ga_a", "ga_v", "ga_l"))
    out = m(a, v, l, "avl")
    assert out.shape == (a.shape[0], 300)
    assert float((out.cpu() - torch.from_numpy(s["ga_out"])).abs().max()) < 1e-5


@pytest.mark.parametrize("N,with_masks", [(1, False), (37, False), (257, True), (1000, True)])
def test_forward_and_gradients_match_oracle(N, with_masks):
    rs = np.random.RandomState(N)
    P = O.formula_weights(SHAPES, seed=11)
    a, v, l = (torch.from_numpy(rs.standard_normal((N, 300)).astype(np.float32)) for _ in range(3))
    gout = torch.from_numpy(rs.standard_normal((N, 300)).astype(np.float32))
    masks = [torch.from_numpy((rs.rand(N, 300) < 0.5).astype(np.uint8)) for _ in range(3)] if with_masks else None

    # oracle (dropout-free restatement applied to the pre-masked inputs; Dropout(0.5) scales kept entries by 2)
    Pc = {k: w.clone().requires_grad_(True) for k, w in P.items()}
    ins = [x.clone().requires_grad_(True) for x in (a, v, l)]
    xs = [x * mk.float() * 2.0 for x, mk in zip(ins, masks)] if with_masks else ins
    ref = O.mm_gated_attention(*xs, Pc)
    ref.backward(gout)

    m = _module(P).train() if with_masks else _module(P).eval()
    din = [x.clone().to(DEV).requires_grad_(True) for x in (a, v, l)]
    out = m(*din, "avl", masks=[mk.to(DEV) for mk in masks] if with_masks else None)
    out.backward(gout.to(DEV))
    assert float((out.detach().cpu() - ref.detach()).abs().max()) < 1e-5

    def rel(x, y):
        return float((x.cpu() - y).norm() / (y.norm() + 1e-12))

    for got, exp in zip(din, ins):
        assert rel(got.grad, exp.grad) < 2e-4
    for name, prm in m.named_parameters():
        assert rel(prm.grad, Pc["gatedatt." + name].grad) < 2e-4, name


def test_train_mode_draws_masks_and_rejects_other_configurations():
    P = O.formula_weights(SHAPES, seed=5)
    m = _module(P).train()
    x = torch.randn(64, 300, device=DEV)
    o1, o2 = m(x, x, x, "avl"), m(x, x, x, "avl")
    assert torch.isfinite(o1).all() and not torch.equal(o1, o2)            # fresh Dropout(0.5) masks per call
    with pytest.raises(NotImplementedError):
        m(x, x, x, "al")
    import mmdfn_b200
    with pytest.raises(NotImplementedError):
        mmdfn_b200.MMGatedAttention(300, 100, att_type='av_bg_fusion').to(DEV)(x, x, x, "avl")
