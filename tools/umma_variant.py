import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
dev="cuda"; torch.manual_seed(0)
def err(ta,tb,M,N,K):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.zeros(M,N,device=dev)
    L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    return float((C.double()-ref).abs().max())
for v in (0,1):
    L.call("mmdfn_gemm_tc_set_variant", v)
    print("variant", v, "NT", err(0,1,128,112,16), "NN K=8", err(0,0,128,112,8), "NN K=16", err(0,0,128,112,16), "NN K=64", err(0,0,256,200,64),
          "TN K=8", err(1,0,128,112,8), "TN K=64", err(1,0,200,112,64))
