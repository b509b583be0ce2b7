"""Times the 2-layer BiGRU entry points (fwd and fwd+bwd) at the bench shapes: text encoder (32 sequences, T = 100)
and party encoder (192 sequences gathered from 9600 rows through a rowmap)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def weights():
    w = []
    for layer in range(2):
        for d in range(2):
            w += [torch.randn(300, 200, device=dev) * 0.05, torch.randn(300, 100, device=dev) * 0.1,
                  torch.randn(300, device=dev) * 0.1, torch.randn(300, device=dev) * 0.1]
    return [t.requires_grad_(True) for t in w]


def run(name, T, nseq, rows, rowmap):
    x = torch.randn(rows, 200, device=dev, requires_grad=True)
    w = weights()
    def fwd():
        return ops.BiGRU2Fn.apply(x, rowmap, T, nseq, None, 1.0, *w)
    y = fwd(); y.sum().backward(); torch.cuda.synchronize()
    for label, fn in (("fwd", lambda: fwd()), ("fwd+bwd", lambda: fwd().backward(torch.ones(T, nseq, 200, device=dev)))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print("%s %-8s %.1f us" % (name, label, e0.elapsed_time(e1) * 1e3 / 20), flush=True)


run("text  (32 seq)", 100, 32, 3200, None)
T, nseq = 100, 192
rm = torch.randint(-1, 9600, (T, nseq), device=dev, dtype=torch.int32)
run("party (192 seq)", T, nseq, 9600, rm)

