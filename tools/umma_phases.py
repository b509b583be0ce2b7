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
    while i + 5 < n - 3:
        a = st[i:i+6]
        nxt = st[i+6] if i + 6 < n - 3 else a[5]
        print(f"  chunk {c:2d}: top@{a[0]:7d} issue_cpasync={a[1]-a[0]:5d} wait_group={a[2]-a[1]:5d} lds+stage_free={a[3]-a[2]:5d} "
              f"split+sts={a[4]-a[3]:5d} fence={a[5]-a[4]:5d} arrive+loop={nxt-a[5]:5d}")
        i += 6; c += 1
    print(f"  mma_done@{st[n-3]}  epilogue+={st[n-2]-st[n-3]}  dealloc+={st[n-1]-st[n-2]} total={st[n-1]}")
run(0,1,38400,300,200)
run(1,0,300,200,19200)
