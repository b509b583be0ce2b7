"""TEST INFRASTRUCTURE (not shipped, not imported by the product): the MFN fusion block (code/model_fusion.py:62-120)
restated in the decomposition the CUDA path will use -- hoisted input GEMMs, three LSTM recurrences, ONE batched
attention / proposal stage over all (t, sequence) rows, a 100-d gated-memory recurrence -- with a hand-written backward
pass (no autograd).  tests/test_oracle_golden.py checks its outputs and every gradient against autograd through
`mmdfn_oracle.mfn_forward`, which is itself pinned to the unmodified reference (tests/golden/mfn.npz).  Purpose: pin the
kernel math (what each stage must save, which products batch into GEMMs) before the kernels exist.

Stages and what they save for the backward:
  S1  pre_m = x_m W_ih^T + b_ih + b_hh                    (T n, 400) per modality      GEMM
  S2  LSTM recurrence per modality                         saves gates i,f,g,o, c_t, h_t
  S3  cStar_t = [c_{t-1}^{lav} | c_t^{lav}]                (T n, 600)                    gather
  S4  A1 = relu(cStar W11^T + b11); att = softmax(A1 W12^T + b12); attended = att * cStar     GEMMs + row softmax
  S5  A2 = relu(attended W21^T + b21); cHat = tanh(A2 W22^T + b22)                             GEMMs
  S6  U1 = attended Wg1a^T + bg1; U2 = attended Wg2a^T + bg2   (gamma*_fc1 split into its attended / memory columns)
  S7  memory recurrence: q_k = relu(U_k[t] + mem Wgkm^T); gamma_k = sigmoid(q_k Wgk2^T + bgk2); mem = gamma1 mem + gamma2 cHat[t]
      saves q1, q2, gamma1, gamma2, mem_{t-1}
  out_t = [h_l | h_a | h_v | mem_t]
"""
from typing import Dict, Tuple

import torch

Tensor = torch.Tensor
MODS = ("l", "a", "v")


def _lin(x, w, b):
    return x @ w.t() + b


def forward(x: Tensor, P: Dict[str, Tensor]) -> Tuple[Tensor, dict]:
    T, n, _ = x.shape
    sv = {"x": x, "T": T, "n": n}
    xs = {"l": x[:, :, :300], "a": x[:, :, 300:600], "v": x[:, :, 600:]}
    # S1 + S2
    c_all = x.new_zeros(T + 1, n, 300)                         # c_all[t + 1] = [c_t^l | c_t^a | c_t^v], c_all[0] = 0
    h_all = x.new_zeros(T, n, 300)
    for mi, m in enumerate(MODS):
        pre = _lin(xs[m].reshape(T * n, 300), P[f"lstm_{m}.weight_ih"], P[f"lstm_{m}.bias_ih"] + P[f"lstm_{m}.bias_hh"]).view(T, n, 400)
        whh = P[f"lstm_{m}.weight_hh"]
        h = x.new_zeros(n, 100)
        c = x.new_zeros(n, 100)
        gates = x.new_zeros(T, n, 400)
        for t in range(T):
            g = pre[t] + h @ whh.t()
            i, f, gg, o = torch.sigmoid(g[:, :100]), torch.sigmoid(g[:, 100:200]), torch.tanh(g[:, 200:300]), torch.sigmoid(g[:, 300:])
            c = f * c + i * gg
            h = o * torch.tanh(c)
            gates[t] = torch.cat([i, f, gg, o], 1)
            c_all[t + 1, :, 100 * mi:100 * mi + 100] = c
            h_all[t, :, 100 * mi:100 * mi + 100] = h
        sv[f"gates_{m}"] = gates
    sv["c_all"], sv["h_all"] = c_all, h_all
    # S3..S6, batched over all T*n rows
    cstar = torch.cat([c_all[:-1], c_all[1:]], dim=2).reshape(T * n, 600)
    A1 = torch.relu(_lin(cstar, P["att1_fc1.weight"], P["att1_fc1.bias"]))
    att = torch.softmax(_lin(A1, P["att1_fc2.weight"], P["att1_fc2.bias"]), dim=1)
    attended = att * cstar
    A2 = torch.relu(_lin(attended, P["att2_fc1.weight"], P["att2_fc1.bias"]))
    chat = torch.tanh(_lin(A2, P["att2_fc2.weight"], P["att2_fc2.bias"]))
    U = {k: _lin(attended, P[f"gamma{k}_fc1.weight"][:, :600], P[f"gamma{k}_fc1.bias"]).view(T, n, 100) for k in (1, 2)}
    sv.update(cstar=cstar, A1=A1, att=att, attended=attended, A2=A2, chat=chat)
    # S7
    mem = x.new_zeros(n, 100)
    mem_prev = x.new_zeros(T, n, 100)
    q = {k: x.new_zeros(T, n, 100) for k in (1, 2)}
    gam = {k: x.new_zeros(T, n, 100) for k in (1, 2)}
    mems = x.new_zeros(T, n, 100)
    chat3 = chat.view(T, n, 100)
    for t in range(T):
        mem_prev[t] = mem
        for k in (1, 2):
            q[k][t] = torch.relu(U[k][t] + mem @ P[f"gamma{k}_fc1.weight"][:, 600:].t())
            gam[k][t] = torch.sigmoid(_lin(q[k][t], P[f"gamma{k}_fc2.weight"], P[f"gamma{k}_fc2.bias"]))
        mem = gam[1][t] * mem + gam[2][t] * chat3[t]
        mems[t] = mem
    sv.update(mem_prev=mem_prev, q=q, gam=gam)
    return torch.cat([h_all, mems], dim=2), sv


def backward(dout: Tensor, P: Dict[str, Tensor], sv: dict) -> Tuple[Tensor, Dict[str, Tensor]]:
    """Hand-written gradients: returns (dx, {parameter name: gradient}).  out_fc1 / out_fc2 get no gradient."""
    T, n, x = sv["T"], sv["n"], sv["x"]
    G: Dict[str, Tensor] = {}
    # ---- S7 backward (reverse time); the per-step products that only feed weight gradients are batched afterwards
    chat3 = sv["chat"].view(T, n, 100)
    dchat = x.new_zeros(T, n, 100)
    dU = {k: x.new_zeros(T, n, 100) for k in (1, 2)}
    dz = {k: x.new_zeros(T, n, 100) for k in (1, 2)}            # d / d(gamma pre-activation)
    carry = x.new_zeros(n, 100)
    for t in range(T - 1, -1, -1):
        dmem = dout[t, :, 300:] + carry
        g1, g2 = sv["gam"][1][t], sv["gam"][2][t]
        dchat[t] = dmem * g2
        carry = dmem * g1
        for k, dg in ((1, dmem * sv["mem_prev"][t]), (2, dmem * chat3[t])):
            gk = sv["gam"][k][t]
            dz[k][t] = dg * gk * (1.0 - gk)
            dq = (dz[k][t] @ P[f"gamma{k}_fc2.weight"]) * (sv["q"][k][t] > 0).float()
            dU[k][t] = dq
            carry = carry + dq @ P[f"gamma{k}_fc1.weight"][:, 600:]
    dattended = x.new_zeros(T * n, 600)
    for k in (1, 2):
        dzk, dUk, qk = dz[k].reshape(T * n, 100), dU[k].reshape(T * n, 100), sv["q"][k].reshape(T * n, 100)
        G[f"gamma{k}_fc2.weight"] = dzk.t() @ qk
        G[f"gamma{k}_fc2.bias"] = dzk.sum(0)
        G[f"gamma{k}_fc1.weight"] = torch.cat([dUk.t() @ sv["attended"], dUk.t() @ sv["mem_prev"].reshape(T * n, 100)], dim=1)
        G[f"gamma{k}_fc1.bias"] = dUk.sum(0)
        dattended = dattended + dUk @ P[f"gamma{k}_fc1.weight"][:, :600]
    # ---- S5 backward
    dP2 = dchat.reshape(T * n, 100) * (1.0 - sv["chat"] ** 2)
    G["att2_fc2.weight"], G["att2_fc2.bias"] = dP2.t() @ sv["A2"], dP2.sum(0)
    dA2 = (dP2 @ P["att2_fc2.weight"]) * (sv["A2"] > 0).float()
    G["att2_fc1.weight"], G["att2_fc1.bias"] = dA2.t() @ sv["attended"], dA2.sum(0)
    dattended = dattended + dA2 @ P["att2_fc1.weight"]
    # ---- S4 backward
    datt = dattended * sv["cstar"]
    dcstar = dattended * sv["att"]
    dS = sv["att"] * (datt - (datt * sv["att"]).sum(1, keepdim=True))
    G["att1_fc2.weight"], G["att1_fc2.bias"] = dS.t() @ sv["A1"], dS.sum(0)
    dA1 = (dS @ P["att1_fc2.weight"]) * (sv["A1"] > 0).float()
    G["att1_fc1.weight"], G["att1_fc1.bias"] = dA1.t() @ sv["cstar"], dA1.sum(0)
    dcstar = (dcstar + dA1 @ P["att1_fc1.weight"]).view(T, n, 600)
    # ---- S3 backward: cStar_t = [c_{t-1} | c_t]
    dc_ext = dcstar[:, :, 300:].clone()                         # w.r.t. c_t
    dc_ext[:-1] += dcstar[1:, :, :300]                          # w.r.t. c_{t-1} of the next step (c_{-1} = 0 is a constant)
    # ---- S2 + S1 backward per modality
    dx = torch.zeros_like(x)
    for mi, m in enumerate(MODS):
        whh, wih = P[f"lstm_{m}.weight_hh"], P[f"lstm_{m}.weight_ih"]
        gates = sv[f"gates_{m}"]
        dpre = x.new_zeros(T, n, 400)
        ch, cc = x.new_zeros(n, 100), x.new_zeros(n, 100)
        for t in range(T - 1, -1, -1):
            i, f, gg, o = gates[t, :, :100], gates[t, :, 100:200], gates[t, :, 200:300], gates[t, :, 300:]
            c_t = sv["c_all"][t + 1, :, 100 * mi:100 * mi + 100]
            c_prev = sv["c_all"][t, :, 100 * mi:100 * mi + 100]
            tc = torch.tanh(c_t)
            dh = dout[t, :, 100 * mi:100 * mi + 100] + ch
            dc = dc_ext[t, :, 100 * mi:100 * mi + 100] + cc + dh * o * (1.0 - tc * tc)
            dpre[t] = torch.cat([dc * gg * i * (1.0 - i), dc * c_prev * f * (1.0 - f), dc * i * (1.0 - gg * gg),
                                 dh * tc * o * (1.0 - o)], 1)
            cc = dc * f
            ch = dpre[t] @ whh
        dpre2 = dpre.reshape(T * n, 400)
        h_prev = torch.cat([x.new_zeros(1, n, 100), sv["h_all"][:-1, :, 100 * mi:100 * mi + 100]], 0).reshape(T * n, 100)
        xm = x[:, :, 300 * mi:300 * mi + 300].reshape(T * n, 300)
        G[f"lstm_{m}.weight_hh"] = dpre2.t() @ h_prev
        G[f"lstm_{m}.weight_ih"] = dpre2.t() @ xm
        G[f"lstm_{m}.bias_ih"] = G[f"lstm_{m}.bias_hh"] = dpre2.sum(0)
        dx[:, :, 300 * mi:300 * mi + 300] = (dpre2 @ wih).view(T, n, 300)
    return dx, G
