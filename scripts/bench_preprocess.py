"""Micro-benchmark of tdrn_preprocess (uint8 BGR frames -> fp32 NCHW network input): CUDA events, achieved GB/s on
algorithmic bytes (frames in + tensor out).  Development aid, not a bench.py number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdrn_b200 import ops

for (b, h, w, s) in ((32, 480, 640, 320), (32, 1080, 1920, 320), (16, 375, 500, 512), (64, 320, 320, 320)):
    f = torch.randint(0, 256, (b, h, w, 3), dtype=torch.uint8, device='cuda')
    out = torch.empty(b, 3, s, s, device='cuda')
    run = lambda: ops.preprocess(f, s, (104, 117, 123), swap_rb=True, out=out)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = b * (h * w * 3 + 3 * s * s * 4)
    print('preprocess b%-2d %4dx%-4d -> %3d  %.4f ms  %.0f GB/s (algorithmic: all source pixels + output)  out-only %.0f GB/s'
          % (b, h, w, s, ms, nbytes / ms / 1e6, b * 3 * s * s * 4 / ms / 1e6))
