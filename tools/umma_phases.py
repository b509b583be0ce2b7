import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
dev="cuda"
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
def run(ta,tb,M,N,K):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.empty(M,N,device=dev)
    for _ in range(3):
        L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_debug", dbg.data_ptr())
    L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_debug", None)
    d = dbg.cpu().tolist(); n = d[127]; t0 = d[0]
    st = [x - t0 for x in d[:n]]
    print(f"--- ta={ta} tb={tb} M={M} N={N} K={K}: {n} stamps; prologue={st[1]}")
    i = 2; c = 0
    while i + 2 < n - 3:
        w, fr, ar = st[i], st[i+1], st[i+2]
        print(f"  chunk {c:2d}: landed@{w:7d} lds+wait_free+={fr-w:5d} convert+fence+arrive+={ar-fr:5d}")
        i += 3; c += 1
    print(f"  mma_done@{st[n-3]}  epilogue+={st[n-2]-st[n-3]}  dealloc+={st[n-1]-st[n-2]} total={st[n-1]}")
run(0,1,38400,300,200)
run(1,0,300,200,19200)
