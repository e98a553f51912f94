"""Raw pinned host->device / device->host bandwidth of the box (what bounds bench.py's e2e): python tools/h2d_probe.py"""
import torch, time
for mb in (19, 38, 154):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device='cuda')
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    up = 20 * n / a.elapsed_time(b) / 1e6
    a.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    dn = 20 * n / a.elapsed_time(b) / 1e6
    print('%4d MiB pinned: H2D %.1f GB/s  D2H %.1f GB/s' % (mb, up, dn), flush=True)
