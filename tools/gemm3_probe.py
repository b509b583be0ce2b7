"""Where does the third-generation GEMM's time go?  Times a few shapes with the profiling switches of umma_gemm3.cu
(no MMAs; narrower column tiles).  GPU box: python tools/gemm3_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
dev = "cuda"
def timed(ta, tb, M, N, K, variant):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.zeros(M, N, device=dev)
    L.call("mmdfn_gemm_tc_set_variant", variant)
    for _ in range(3):
        L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N, None, 0, L.stream())
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_variant", 0)
    return e0.elapsed_time(e1) * 1e3 / 20
for name, ta, tb, M, N, K in (("NT big", 0, 1, 153600, 300, 200), ("NT big K=800", 0, 1, 153600, 300, 800), ("NT gru in-gemm", 0, 1, 19200, 300, 200), ("NT one wave", 0, 1, 18944, 160, 200),
                              ("NT one wave K=800", 0, 1, 18944, 160, 800), ("NT one wave K=16", 0, 1, 18944, 160, 16)):
    r = {v: timed(ta, tb, M, N, K, v) for v in (3, 31, 37, 2)}
    print(f"{name:20s} M={M} N={N} K={K}: gen3 {r[3]:7.1f} us | no-MMA {r[31]:7.1f} | no B loads {r[37]:7.1f} | gen2 {r[2]:7.1f}", flush=True)
