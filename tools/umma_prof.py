"""Launch a few representative mmdfn_gemm_tc problems (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
dev = "cuda"
def run(ta, tb, M, N, K, reps=3):
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.empty(M, N, device=dev)
    for _ in range(reps):
        L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
run(0, 1, 38400, 300, 200)
run(1, 0, 300, 200, 19200)
