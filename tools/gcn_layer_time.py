"""Duration of the fused graph-conv layer kernel alone (bench.roofline_graph_conv protocol) per kernel generation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mmdfn_b200 import _lib as L
dev = torch.device("cuda", 0)
for variant in [int(v) for v in (sys.argv[1:] or ["0", "2"])]:
    print("=== variant %d ===" % variant, flush=True)
    L.call("mmdfn_gcn_layer_set_variant", variant)
    for nd in (32, 64, 148, 256, 512):
        r = bench.roofline_graph_conv(dev, nd)
        print("%4d dialogues x 100: %.2f us/launch  achieved %.0f GB/s  frac %.3f   (copy kernel %.2f us, frac %.3f)" % (
            nd, r["us_per_launch"], r["achieved"], r["frac"], r["same_bytes_copy_kernel"]["us_per_launch"], r["same_bytes_copy_kernel"]["frac"]), flush=True)
L.call("mmdfn_gcn_layer_set_variant", 0)
