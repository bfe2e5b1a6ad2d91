"""Probe: pinned host <-> device copy rate for the bench's per-step buffers (39.3 MB in, 2.7 MB out)."""
import torch
x = torch.randn(32, 3, 320, 320).pin_memory()
d = torch.empty_like(x, device='cuda')
o = torch.empty(32, 21, 200, 5, device='cuda')
ho = torch.empty(32, 21, 200, 5).pin_memory()
for name, fn, nb in (('H2D 39.3 MB', lambda: d.copy_(x, non_blocking=True), x.numel() * 4),
                     ('D2H 2.7 MB', lambda: ho.copy_(o, non_blocking=True), o.numel() * 4)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%s: %.3f ms  %.1f GB/s' % (name, ms, nb / ms / 1e6))
