"""Run mmdfn_gemm_tc on ONE shape a few times (ncu target): python tools/gemm_one.py ta tb M N K [variant]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
ta, tb, M, N, K = (int(x) for x in sys.argv[1:6])
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 0
torch.manual_seed(0)
A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((N, K) if tb else (K, N), device="cuda")
C = torch.zeros(M, N, device="cuda")
L.call("mmdfn_gemm_tc_set_variant", variant)
for _ in range(4):
    L.call("mmdfn_gemm_tc", ta, tb, M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
torch.cuda.synchronize()
ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
print("max err", float((C.double() - ref).abs().max()))
