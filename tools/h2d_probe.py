"""Pinned host -> device copy bandwidth on this box (the e2e bench line's H2D leg)."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (1, 8, 21, 64, 256):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device=dev)
    s = torch.cuda.Stream()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb:4d} MB pinned: {ms:7.3f} ms  {mb / 1024 / (ms / 1e3):6.1f} GB/s", flush=True)
    h2 = torch.empty(4, dtype=torch.float32).pin_memory()
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H {mb:4d} MB pinned: {ms:7.3f} ms  {mb / 1024 / (ms / 1e3):6.1f} GB/s", flush=True)
