"""clock64 phase stamps of CTA 0 of the tcgen05 aggregate kernel (mmdfn_adj_spmm_set_debug)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L, ops

dev = "cuda"
dbg = torch.zeros(64, dtype=torch.int64, device=dev)
for nd in (1, 32, 256):
    lengths = [100] * nd
    geom = ops.DialogGeom(lengths, dev)
    N = geom.N
    blk = torch.rand(geom.nblk, device=dev) / 100
    dg = torch.rand(3, N, device=dev) / 100
    x = torch.randn(3 * N, 100, device=dev)
    y = torch.empty(3 * N, 100, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for cold in (0, 1):
        for _ in range(2):
            L.call("mmdfn_adj_spmm", *geom.args(), L.ptr(blk), L.ptr(dg), L.ptr(x), 100, L.ptr(y), L.stream())
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        L.call("mmdfn_adj_spmm_set_debug", dbg.data_ptr())
        L.call("mmdfn_adj_spmm", *geom.args(), L.ptr(blk), L.ptr(dg), L.ptr(x), 100, L.ptr(y), L.stream())
        torch.cuda.synchronize()
        L.call("mmdfn_adj_spmm_set_debug", None)
        d = dbg.cpu().tolist()
        n = d[63]
        st = [v - d[0] for v in d[:n]]
        print("dialogues %d %s: %d stamps" % (nd, "cold" if cold else "warm", n))
        print("  loads issued @%d, set-up done @%d" % (st[1], st[2]))
        i, c = 3, 0
        while i + 2 < n - 3:
            print("  chunk %d: z landed @%6d  stage free +%5d  converted +%5d" % (c, st[i], st[i + 1] - st[i], st[i + 2] - st[i + 1]))
            i += 3; c += 1
        print("  mma done @%d  tile in smem +%d  stored +%d  total %d cycles" % (st[n - 3], st[n - 2] - st[n - 3], st[n - 1] - st[n - 2], st[n - 1]))
