"""SUPERSEDED by scripts/umma_rate_bg.py: the MMA rows of this probe are limited by its own issue loop (run-time `i % nacc`
per group of four MMAs), not by the hardware -- see profiles/probe_umma_rate2.txt.  Original question:
where does the 73-cycle floor of narrow-N tcgen05.mma come from (tdrn_debug_umma_rate2)?  Cycles per K = 16
instruction (clock64) and chip TFLOP/s (CUDA events) for A from shared memory (SS), A from tensor memory (TS),
tcgen05.cp alone, cp + TS interleaved, and SS with cta_group::2 (M = 256 over an SM pair).

    python scripts/umma_rate2.py [modes...]       # default: 0 1 2 3 5 6 4 (4 = cta_group::2 runs last)
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import _lib
L = _lib.probe_lib()
NAMES = {0: 'SS', 1: 'TS (A in TMEM)', 2: 'cp 128x256b only', 3: 'cp + TS', 4: 'SS cta_group::2 (M=256)',
         5: 'SS, descs per 4 MMAs', 6: 'SS, descs per MMA'}
grid, iters = 148, 20000
modes = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 5, 6, 4]
cyc = torch.zeros(grid, dtype=torch.int64, device='cuda')
for mode in modes:
    for n in (64, 128, 256):
        for nacc in ((1,) if mode == 2 else (1, 2)):
            if nacc * n > 448:
                continue
            cyc.zero_()
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(L.tdrn_debug_umma_rate2(ctypes.c_void_p(cyc.data_ptr()), grid, n, iters, nacc, mode), 'rate2')
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            units = grid // 2 if mode == 4 else grid              # issuing CTAs
            m = 256 if mode == 4 else 128
            c = cyc[:units].float().mean().item() / (iters * 4)
            tf = 0.0 if mode == 2 else units * iters * 4 * 2.0 * m * n * 16 / (ms * 1e-3) / 1e12
            print('%-24s N=%3d nacc=%d: %6.1f cycles per instruction (N/2 = %3d)  %5.0f TFLOP/s chip  (%.2f ms)' %
                  (NAMES[mode], n, nacc, c, n // 2, tf, ms), flush=True)
