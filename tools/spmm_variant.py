"""A/B of the aggregate kernels (variant 0: tcgen05 for L <= 128 else FFMA, 1: FFMA only, 2: experimental any-length
tcgen05 kernel): max error against an fp64 dense product and time per launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L, ops

dev = "cuda"


def dense_ref(blk, dg, x, lengths):
    N = sum(lengths)
    y = torch.zeros_like(x, dtype=torch.float64)
    xd = x.double()
    off, bo = 0, 0
    pair = {(0, 1): 0, (0, 2): 1, (1, 2): 2}
    for Lb in lengths:
        for m in range(3):
            A = blk[bo + m * Lb * Lb: bo + (m + 1) * Lb * Lb].view(Lb, Lb).double()
            y[m * N + off:m * N + off + Lb] += A @ xd[m * N + off:m * N + off + Lb]
            for n in range(3):
                if n != m:
                    d = dg[pair[(min(m, n), max(m, n))], off:off + Lb].double()
                    y[m * N + off:m * N + off + Lb] += d[:, None] * xd[n * N + off:n * N + off + Lb]
        off += Lb
        bo += 3 * Lb * Lb
    return y


def run(lengths, variant, reps=0):
    torch.manual_seed(0)
    geom = ops.DialogGeom(lengths, dev)
    N = geom.N
    blk = torch.rand(geom.nblk, device=dev) / max(lengths)
    dg = torch.rand(3, N, device=dev) / max(lengths)
    x = torch.randn(3 * N, 100, device=dev)
    y = torch.full((3 * N, 100), float("nan"), device=dev)
    L.call("mmdfn_adj_spmm_set_variant", variant)
    L.call("mmdfn_adj_spmm", *geom.args(), L.ptr(blk), L.ptr(dg), L.ptr(x), 100, L.ptr(y), L.stream())
    torch.cuda.synchronize()
    err = float((y.double() - dense_ref(blk, dg, x, lengths)).abs().max())
    us = None
    if reps:
        copies = max(2, int(300e6 // (4 * (6 * N * 100 + geom.nblk))) + 1)
        bl = [blk.clone() for _ in range(copies)]; xs = [x.clone() for _ in range(copies)]; ys = [torch.empty_like(y) for _ in range(copies)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(copies):
            L.call("mmdfn_adj_spmm", *geom.args(), L.ptr(bl[i]), L.ptr(dg), L.ptr(xs[i]), 100, L.ptr(ys[i]), L.stream())
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            j = i % copies
            L.call("mmdfn_adj_spmm", *geom.args(), L.ptr(bl[j]), L.ptr(dg), L.ptr(xs[j]), 100, L.ptr(ys[j]), L.stream())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
    L.call("mmdfn_adj_spmm_set_variant", 0)
    return err, us


if __name__ == "__main__":
    # variant 2 = the experimental any-length tcgen05 kernel (spmm_tc_long.cu): first the correctness-only cases, then timing
    cases = ([100], [5, 3, 7], [128, 1, 37, 64, 99, 33], [129], [200, 17, 131], [500], [257, 3, 128, 255],
             [100] * 32, [100] * 256, [96] * 32, [128] * 32, [64] * 32, [200] * 96, [500] * 24)
    for lengths in cases:
        for v in (0, 1, 2):
            err, us = run(lengths, v, reps=200 if len(lengths) >= 24 else 0)
            print("lengths %s x%d variant %d: max err %.3g  %s" % (lengths[:6], len(lengths), v, err, "" if us is None else "%.2f us" % us), flush=True)
