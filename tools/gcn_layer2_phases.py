"""clock64 stamps of CTA 0 of the second-generation fused graph-conv layer kernel (gcn_layer2.cu), per tile and role."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
ROW = ["start", "mma_a seen", "cross landed", "hop done", "xy released", "mma_b seen", "drained", "R/q landed", "epilogue math", "row barrier", "stored"]
CONV = {12: "start", 13: "A raw landed", 14: "prev mma_b seen", 15: "A in TMEM", 16: "z chunks done"}
MMA = {18: "a_ready", 19: "d_free", 20: "A chunk0 full", 21: "A last full", 22: "t_ready", 23: "B chunk0 full", 24: "B last full"}


def stamps(lengths, fwd=True):
    geom = ops.DialogGeom(lengths, dev)
    N = geom.N
    n3 = 3 * N
    blk = torch.rand(geom.nblk, device=dev) / max(lengths)
    dg = torch.rand(3, N, device=dev) / max(lengths)
    z, r, q = (torch.randn(n3, 100, device=dev) for _ in range(3))
    y = torch.empty(n3, 100, device=dev)
    mk = (torch.rand(n3, 100, device=dev) > 0.4).to(torch.uint8)
    fl = torch.empty(n3, 100, device=dev, dtype=torch.uint8)
    W = [torch.randn(200, 100, device=dev) * 0.1]
    img_n = L.query("mmdfn_gcn_layer_img_floats")
    mtop, mbot = torch.empty(100, 100, device=dev), torch.empty(100, 100, device=dev)
    img_f, img_b = torch.empty(img_n, device=dev), torch.empty(img_n, device=dev)
    L.call("mmdfn_gcn_layer_prep", 1, L.ptr_table(W), 0.5, 0.2, L.ptr(mtop), L.ptr(mbot), L.ptr(img_f), L.ptr(img_b), L.stream())
    dbg = torch.zeros(512, dtype=torch.int64, device=dev)
    for rep in range(3):
        dbg.zero_()
        L.call("mmdfn_gcn_layer_set_debug", dbg.data_ptr())
        L.call("mmdfn_gcn_layer_fwd", *geom.args(), L.ptr(blk), L.ptr(dg), L.ptr(z), L.ptr(img_f), L.ptr(r), 100, L.ptr(q), L.ptr(mk, torch.uint8), 1.0 / 0.6,
               L.ptr(fl, torch.uint8), L.ptr(y), 100, L.stream())
        L.call("mmdfn_gcn_layer_set_debug", None)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist()
    t0 = min(x for x in d if x > 0)
    print("lengths %s x%d (cycles since the CTA's first stamp)" % (lengths[:3], len(lengths)))
    for it in range(8):
        b = d[32 * it:32 * it + 32]
        if not any(b):
            break
        print("  tile %d  row : %s" % (it, "  ".join("%s %d" % (ROW[i], b[i] - t0) for i in range(11) if b[i])))
        print("          conv: %s" % "  ".join("%s %d" % (CONV[i], b[i] - t0) for i in sorted(CONV) if b[i]))
        print("          mma : %s" % "  ".join("%s %d" % (MMA[i], b[i] - t0) for i in sorted(MMA) if b[i]))
    f = d[256:]
    if any(f):
        print("  tile 1 converter, z chunk c: splits | loads issued | stage free | stores | fence | arrive   (cycles since the chunk's start; start since first stamp)")
        for c in range(7):
            b = f[8 * c:8 * c + 7]
            if b[0]:
                print("     c=%d start %6d: %s" % (c, b[0] - t0, " | ".join("%5d" % (b[i] - b[0]) for i in range(1, 7))))
        print("  tile 1 converter, A_hat block jj: loaded->stored ", " ".join("%d:%d" % (f[64 + 2 * j] - t0, f[65 + 2 * j] - f[64 + 2 * j]) for j in range(8) if f[64 + 2 * j]), " wait::st done", f[80] - t0)
        print("  tile 1 hop blocks done at", " ".join("%d" % (f[96 + j] - t0) for j in range(7) if f[96 + j]), " wait::st", f[104] - t0)


if __name__ == "__main__":
    L.call("mmdfn_gcn_layer_set_variant", 0)
    stamps([100] * 32)
    stamps([100] * 256)
