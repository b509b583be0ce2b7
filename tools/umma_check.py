"""Standalone check of mmdfn_gemm_tc (tcgen05 3xTF32 GEMM) against fp64, printing the error per shape and mode
next to the FFMA kernel's; also times both, the three kernel generations and torch.matmul (cuBLAS) on the same shapes.  Usage (GPU box): python tools/umma_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"

def run(name, ta, tb, M, N, K, bias=False, act=0, beta=0.0, alpha=1.0):
    A = torch.randn((K, M) if ta else (M, K), device=dev)
    B = torch.randn((N, K) if tb else (K, N), device=dev)
    C0 = torch.randn(M, N, device=dev)
    bi = torch.randn(N, device=dev) if bias else None
    opA = A.double().t() if ta else A.double()
    opB = B.double().t() if tb else B.double()
    ref = alpha * (opA @ opB) + beta * C0.double() + (bi.double() if bias else 0)
    if act: ref = torch.relu(ref)
    out = {}
    for fn in ("mmdfn_gemm_tc", "mmdfn_gemm"):
        C = C0.clone()
        L.call(fn, int(ta), int(tb), M, N, K, alpha, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], beta, L.ptr(C), N,
               L.ptr(bi) if bias else None, act, L.stream())
        torch.cuda.synchronize()
        out[fn] = float((C.double() - ref).abs().max())
        # timing
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.call(fn, int(ta), int(tb), M, N, K, alpha, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N,
                   L.ptr(bi) if bias else None, act, L.stream())
        e1.record(); torch.cuda.synchronize()
        out[fn + "_us"] = e0.elapsed_time(e1) * 1e3 / 5
    gf = 2.0 * M * N * K / 1e9
    print(f"{name:28s} M={M:6d} N={N:4d} K={K:6d} err_tc={out['mmdfn_gemm_tc']:.2e} err_ffma={out['mmdfn_gemm']:.2e} "
          f"tc={out['mmdfn_gemm_tc_us']:8.1f}us ({gf/out['mmdfn_gemm_tc_us']*1e3:7.1f} TF/s) ffma={out['mmdfn_gemm_us']:8.1f}us "
          f"({gf/out['mmdfn_gemm_us']*1e3:6.1f} TF/s)", flush=True)
    return out["mmdfn_gemm_tc"]

print(torch.cuda.get_device_name(0), flush=True)
run("NT tiny", 0, 1, 128, 112, 32)
run("NT one tile K=8", 0, 1, 128, 16, 8)
run("NT ragged", 0, 1, 130, 100, 100, bias=True, act=1, beta=0.5, alpha=0.7)
run("NT proj text", 0, 1, 3200, 200, 100, bias=True)
run("NT proj audio", 0, 1, 3200, 200, 512, bias=True)
run("NT proj visual", 0, 1, 3200, 200, 1024, bias=True)
run("NT proj iemocap audio", 0, 1, 3520, 200, 1582, bias=True)
run("NT proj iemocap visual", 0, 1, 3520, 200, 342, bias=True)
run("NT gru in-gemm", 0, 1, 19200, 300, 200, bias=True)
run("NT lstm gates", 0, 1, 9600, 400, 100, bias=True)
run("NT big", 0, 1, 153600, 300, 200, bias=True)
run("NN dx", 0, 0, 19200, 200, 300)
run("NN conv W", 0, 0, 9600, 100, 100, beta=1.0)
run("NN ragged", 0, 0, 77, 45, 19)
run("TN dW_ih splitK", 1, 0, 300, 200, 19200)
run("TN dW small", 1, 0, 100, 100, 9600, beta=1.0)
run("TN dW proj", 1, 0, 200, 1024, 3200)
run("TN ragged", 1, 0, 77, 45, 190)
print("=== generation 1 (cp.async staging ring) / 2 (register / TMA operand paths) / 3 (A operand in tensor memory, wide CTA): us per launch")
def timed(ta, tb, M, N, K, variant, beta=0.0):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.zeros(M, N, device=dev)
    L.call("mmdfn_gemm_tc_set_variant", variant)
    for _ in range(3):
        L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], beta, L.ptr(C), N, None, 0, L.stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            L.call("mmdfn_gemm_tc", int(ta), int(tb), M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], beta, L.ptr(C), N, None, 0, L.stream())
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_variant", 0)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    return e0.elapsed_time(e1) * 1e3 / 20
def cublas(ta, tb, M, N, K, tf32):
    """torch.matmul on the same operands and layouts (cuBLAS / cuBLASLt), graph-replayed like the kernels above"""
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev); C = torch.zeros(M, N, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = tf32
    opA = A.t() if ta else A; opB = B.t() if tb else B
    for _ in range(3):
        torch.matmul(opA, opB, out=C)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            torch.matmul(opA, opB, out=C)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    return e0.elapsed_time(e1) * 1e3 / 20
for name, ta, tb, M, N, K in (("NT proj audio", 0, 1, 3200, 200, 512), ("NT proj visual", 0, 1, 3200, 200, 1024), ("NT gru in-gemm party", 0, 1, 9600, 600, 200),
                              ("NT gru in-gemm", 0, 1, 19200, 300, 200), ("NT lstm gates", 0, 1, 9600, 400, 100), ("NT big", 0, 1, 153600, 300, 200),
                              ("NN dx", 0, 0, 19200, 200, 300), ("NN dgates W", 0, 0, 9600, 100, 400), ("NN R_all", 0, 0, 9600, 200, 100),
                              ("TN dW_ih splitK", 1, 0, 300, 200, 19200), ("TN dW gates", 1, 0, 400, 100, 9600), ("TN dW proj", 1, 0, 200, 1024, 3200),
                              ("NT proj text", 0, 1, 3200, 200, 100), ("NN dx party l0", 0, 0, 9600, 200, 600), ("TN dW_ih l1 both", 1, 0, 600, 200, 19200)):
    t1, t2, t3 = timed(ta, tb, M, N, K, 1), timed(ta, tb, M, N, K, 2), timed(ta, tb, M, N, K, 3)
    tc32, tctf = cublas(ta, tb, M, N, K, False), cublas(ta, tb, M, N, K, True)
    gf = 2.0 * M * N * K / 1e9
    print(f"{name:24s} M={M:6d} N={N:4d} K={K:6d}  gen1 {t1:7.1f} us ({gf/t1*1e3:6.1f} TF/s)   gen2 {t2:7.1f} us ({gf/t2*1e3:6.1f} TF/s)   "
          f"gen3 {t3:7.1f} us ({gf/t3*1e3:6.1f} TF/s)   gen3 vs best x{min(t1, t2)/t3:.2f}   | torch.matmul fp32 (cuBLAS) {tc32:7.1f} us ({gf/tc32*1e3:6.1f} TF/s), "
          f"allow_tf32 (1xTF32, NOT fp32-accurate) {tctf:7.1f} us", flush=True)
for bn in (() if "--quick" in sys.argv else (112, 160, 224)):
    L.call("mmdfn_gemm_tc_set_variant", bn)
    print(f"--- forced BN={bn}")
    run("NT gru in-gemm", 0, 1, 19200, 300, 200, bias=True)
    run("NT ragged", 0, 1, 130, 100, 100, bias=True, act=1, beta=0.5, alpha=0.7)
    run("NT big", 0, 1, 153600, 300, 200, bias=True)
    run("NN dx", 0, 0, 19200, 200, 300)
    run("NN ragged", 0, 0, 77, 45, 19)
    run("TN dW_ih splitK", 1, 0, 300, 200, 19200)
    run("TN ragged", 1, 0, 277, 245, 190)
L.call("mmdfn_gemm_tc_set_variant", 0)
print("done")
