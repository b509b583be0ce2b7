"""clock64 phase stamps of CTA (0,0) of the fused graph-conv layer kernel (mmdfn_gcn_layer_set_debug) and its duration
alone (graph-replayed back-to-back launches, rotating operands) at several shard sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mmdfn_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
NAMES = ["entry", "setup", "A converted", "A mma done", "T tile", "B converted", "B mma done", "stored"]


def stamps(lengths, detail=False):
    geom = ops.DialogGeom(lengths, dev)
    N = geom.N
    n3 = 3 * N
    blk = torch.rand(geom.nblk, device=dev) / max(lengths)
    dg = torch.rand(3, N, device=dev) / max(lengths)
    z, r, q = (torch.randn(n3, 100, device=dev) for _ in range(3))
    y = torch.empty(n3, 100, device=dev)
    fl = torch.empty(n3, 100, device=dev, dtype=torch.uint8)
    W = [torch.randn(200, 100, device=dev) * 0.1]
    img_n = L.query("mmdfn_gcn_layer_img_floats")
    mtop, mbot = torch.empty(100, 100, device=dev), torch.empty(100, 100, device=dev)
    img_f, img_b = torch.empty(img_n, device=dev), torch.empty(img_n, device=dev)
    L.call("mmdfn_gcn_layer_prep", 1, L.ptr_table(W), 0.5, 0.2, L.ptr(mtop), L.ptr(mbot), L.ptr(img_f), L.ptr(img_b), L.stream())
    dbg = torch.zeros(256, dtype=torch.int64, device=dev)
    for rep in range(3):
        L.call("mmdfn_gcn_layer_set_debug", dbg.data_ptr())
        L.call("mmdfn_gcn_layer_fwd", *geom.args(), L.ptr(blk), L.ptr(dg), L.ptr(z), L.ptr(img_f), L.ptr(r), 100, L.ptr(q), None, 1.0,
               L.ptr(fl, torch.uint8), L.ptr(y), 100, L.stream())
        L.call("mmdfn_gcn_layer_set_debug", None)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    n = d[63]
    print("lengths %s x%d: %s" % (lengths[:3], len(lengths), "  ".join("%s +%d" % (NAMES[i], d[i] - d[i - 1]) for i in range(1, n))),
          " total %d cycles" % (d[n - 1] - d[0]), flush=True)
    if detail:
        t0 = d[0]
        ng = min(60, -(-lengths[0] // 8) + 13) if L.query('mmdfn_gcn_layer_img_floats') < 25000 else min(60, -(-lengths[0] // 16) + 7)
        print("   phase A chunk 3: A stores +%d  B pieces +%d  fence +%d  syncwarp +%d  arrive +%d" % tuple(d[241 + i] - d[240 + i] for i in range(5)))
        print("   phase B chunk 3: A stores +%d  W stores +%d  fence +%d  syncwarp+arrive +%d" % tuple(d[249 + i] - d[248 + i] for i in range(4)))
        print("   k-step: issuer-full | conv-reached-wait | conv-stage-free   (cycles since entry)")
        for g in range(ng):
            print("   %2d: %6d | %6d | %6d" % (g, d[64 + g] - t0, d[192 + g] - t0, d[128 + g] - t0))


if __name__ == "__main__":
  for variant in (0, 1):
    print("=== variant %d (%s) ===" % (variant, "KC 16 / 2 stages" if variant == 0 else "KC 8 / 3 stages"), flush=True)
    L.call("mmdfn_gcn_layer_set_variant", variant)
    stamps([100] * 32, detail=True)
    for lengths in ([100] * 256, [50] * 64, [128] * 32, [500] * 24):
        stamps(lengths)
    for nd in (32, 64, 128, 256, 512):
        r = bench.roofline_graph_conv(dev, nd)
        print("%4d dialogues x 100: %.2f us/launch  achieved %.0f GB/s  frac %.3f   (copy kernel %.2f us, frac %.3f)" % (
            nd, r["us_per_launch"], r["achieved"], r["frac"], r["same_bytes_copy_kernel"]["us_per_launch"], r["same_bytes_copy_kernel"]["frac"]), flush=True)
