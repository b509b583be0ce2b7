"""A operand from tensor memory (tcgen05.mma [d], [a_tmem], b_desc): read-back check of the lane/column convention."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
for a_col in (16, 128, 200):
    out = torch.full((128, 16), -1.0, device="cuda")
    L.call("mmdfn_umma_probe_ta", L.ptr(out), a_col, L.stream()); torch.cuda.synchronize()
    want = (16 * torch.arange(128).view(128, 1) + torch.arange(16).view(1, 16)).float()
    o = out.cpu()
    print("a_col", a_col, "exact" if torch.equal(o, want) else "MISMATCH")
    if not torch.equal(o, want):
        print(o[:4].int()); print(o[60:68].int()); print(o[124:].int())
