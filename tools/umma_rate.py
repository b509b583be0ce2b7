"""Steady-state cost of one tcgen05.mma (128 x N x 32 bytes of K) per kind / N / A-operand source / number of
independent accumulators the instructions rotate over."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
out = torch.zeros(2, dtype=torch.int64, device="cuda")
reps = 64
for kind, kn in ((0, "tf32 K=8 "), (1, "bf16 K=16")):
    for a_tmem in (0, 1):
        for N, nacc in ((64, 1), (64, 2), (64, 4), (112, 1), (128, 1), (224, 1), (256, 1)):
            for _ in range(2):
                L.call("mmdfn_umma_rate", out.data_ptr(), kind, N, a_tmem, reps, nacc, L.stream()); torch.cuda.synchronize()
            a, b = out.cpu().tolist()
            print("%s  A from %s  N=%3d  issuers=%d: %6.1f cycles per MMA (all issuers together)  (%d for %d, %d for %d)" % (kn, "TMEM" if a_tmem else "smem", N, nacc, (b - a) / reps / nacc, a, reps, b, 2 * reps), flush=True)
