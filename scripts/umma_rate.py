"""SUPERSEDED by scripts/umma_rate_bg.py (this probe's 73-cycle rows are its own issue loop, see profiles/probe_umma_rate.txt).
tcgen05.mma sustained rate for narrow tiles (tdrn_debug_umma_rate): cycles per K=16 MMA (clock64) and
TFLOP/s over the whole chip (CUDA events) for N in 64/128/256, 1 or 2 accumulators, aligned or row-shifted A."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import _lib
L = _lib.probe_lib()
grid, iters = 148, 20000
cyc = torch.zeros(grid, dtype=torch.int64, device='cuda')
for n in (64, 128, 256):
    for nacc in (1, 2):
        for shift in (0, 11):
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(L.tdrn_debug_umma_rate(ctypes.c_void_p(cyc.data_ptr()), grid, n, iters, nacc, shift), 'rate')
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            c = cyc.float().mean().item() / (iters * 4)
            tf = grid * iters * 4 * 2.0 * 128 * n * 16 / (ms * 1e-3) / 1e12
            print('N=%3d nacc=%d a_shift=%2d: %.1f cycles per MMA (ideal %.0f)  %.0f TFLOP/s chip  (%.2f ms, %.0f MHz effective)' %
                  (n, nacc, shift, c, n / 2, tf, ms, cyc.float().mean().item() / (ms * 1e3)))
