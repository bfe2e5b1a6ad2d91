"""tcgen05.mma issue rate next to other shared-memory traffic (tdrn_debug_umma_rate_bg): does the shared-memory
bandwidth the tensor core needs for its operands (A 4 KB + B N x 32 B per K = 16 instruction) explain why the N = 128
layers run at ~1.35x the 71-cycle floor, and does cta_group::2 (each SM reads half of B) relieve it?
Not run yet (written after the round-1 GPU budget was spent): first experiment of the next round.

    python scripts/umma_rate_bg.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import _lib
L = _lib.lib()
grid, iters = 148, 20000
cyc = torch.zeros(grid, dtype=torch.int64, device='cuda')
bgb = torch.zeros(grid, dtype=torch.int64, device='cuda')
for n in (64, 128, 256):
    for cg in (1, 2):
        for kind in (0, 1):
            for warps, gap in ((0, 0), (1, 256), (1, 0), (2, 0), (3, 0)):
                if warps == 0 and kind == 1:
                    continue
                cyc.zero_(); bgb.zero_()
                for rep in range(2):
                    _lib.check(L.tdrn_debug_umma_rate_bg(ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(bgb.data_ptr()), grid, n, iters,
                                                         cg, warps, kind, gap), 'rate_bg')
                    torch.cuda.synchronize()
                units = grid // cg
                c = cyc[:units].float().mean().item()
                per = c / (iters * 4)
                bg = bgb.float().mean().item() / max(c, 1.0)
                tc = (4096 + n * 32 // cg) / per                 # operand bytes each SM's tensor core reads per clock
                print('N=%3d cta_group::%d  background %s x%d warps gap %3d: %6.1f cycles per MMA  tensor-core reads %5.1f B/clk/SM  '
                      'background %5.1f B/clk/SM' % (n, cg, ('loads ', 'stores')[kind], warps, gap, per, tc, bg), flush=True)
