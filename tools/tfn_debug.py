import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
torch.manual_seed(0)
W = 1030301
for rows in (5, 256):
    dy = torch.randn(rows, 300, device="cuda")
    F = torch.randn(rows, W + 3, device="cuda")
    C = torch.full((300, W), 7.0, device="cuda")
    L.call("mmdfn_gemm", 1, 0, 300, W, rows, 1.0, L.ptr(dy), 300, L.ptr(F), W + 3, 0.0, L.ptr(C), W, None, 0, L.stream())
    torch.cuda.synchronize()
    ref = dy.t().double() @ F[:, :W].double()
    d = (C.double() - ref).abs()
    print("TN rows", rows, "max err", float(d.max()), "n bad", int((d > 1e-3).sum()))
    bad = (d > 1e-3).nonzero()
    if len(bad):
        print(" first bad", bad[:5].tolist(), "last bad", bad[-5:].tolist(), "cols mod 192", sorted(set((bad[:, 1] % 192).tolist()))[:20], "rows", sorted(set(bad[:,0].tolist()))[:10])
    # NN: dF = dy (rows,300) @ W1 (300, W)
    W1 = torch.randn(300, W, device="cuda") * 0.01
    out = torch.full((rows, W + 3), 7.0, device="cuda")
    L.call("mmdfn_gemm", 0, 0, rows, W, 300, 1.0, L.ptr(dy), 300, L.ptr(W1), W, 0.0, L.ptr(out), W + 3, None, 0, L.stream())
    torch.cuda.synchronize()
    ref = dy.double() @ W1.double()
    d = (out[:, :W].double() - ref).abs()
    print("NN rows", rows, "max err", float(d.max()), "n bad", int((d > 1e-3).sum()))
    # NT: y = F (rows, W) @ W1^T
    y = torch.zeros(rows, 300, device="cuda")
    L.call("mmdfn_gemm", 0, 1, rows, 300, W, 1.0, L.ptr(F), W + 3, L.ptr(W1), W, 1.0, L.ptr(y), 300, None, 0, L.stream())
    torch.cuda.synchronize()
    ref = F[:, :W].double() @ W1.double().t()
    print("NT rows", rows, "max err", float((y.double() - ref).abs().max()), "ref max", float(ref.abs().max()))
