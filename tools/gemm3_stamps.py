"""clock64 stamps inside umma_gemm3_kernel (CTA 0, converter warp 0 = group 0, the chunks 0, 4, 8, ... it owns)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
M, N, K = 18944, 160, 800
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
for variant in (3,):
    L.call("mmdfn_gemm_tc_set_variant", variant)
    for _ in range(2):
        L.call("mmdfn_gemm_tc", 0, 1, M, N, K, 1.0, L.ptr(A), K, L.ptr(B), K, 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_debug", L.ptr(buf, torch.int64))
    L.call("mmdfn_gemm_tc", 0, 1, M, N, K, 1.0, L.ptr(A), K, L.ptr(B), K, 0.0, L.ptr(C), N, None, 0, L.stream())
    torch.cuda.synchronize()
    L.call("mmdfn_gemm_tc_set_debug", None)
    t = buf.cpu().tolist(); t0 = t[0]
    print("variant", variant, "(stamps: start | before raw wait | raw landed | A st issued | B stored | prefetch issued | st waited | arrived)")
    for it in range(8):
        print("  chunk %2d:" % (4 * it), " ".join("%6d" % (t[8 * it + j] - t0) for j in range(8)))
L.call("mmdfn_gemm_tc_set_variant", 0)
