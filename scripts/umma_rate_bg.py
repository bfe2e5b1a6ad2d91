"""tcgen05.mma issue rate with the descriptors resident in uniform registers (tdrn_debug_umma_rate_bg), next to other
shared-memory traffic, with cta_group 1 or 2, and with the row-shifted / 1280-byte-strided A views of the halo-tile kernels.

First run (r01f, `profiles/probe_umma_rate_bg.txt`, aligned A only): 48.0 cycles per N = 64 instruction and 64.0 per N = 128
-- the tensor core reads its operands at exactly 128 B/clk -- so the "71-73 cycle floor" of scripts/umma_rate.py /
umma_rate2.py was their own issue loop (a run-time `i % nacc` per iteration), not the hardware.  The A-view sweep at the
end of this script has NOT run yet: it asks whether the production halo kernels' 84 / 96-100 cycles per MMA (N = 64 / 128,
epilogue skipped) are the shifted A views reading at half rate (4 KB at 64 B/clk + B would give 80 / 96).

    python scripts/umma_rate_bg.py [views]        # "views": only the A-view sweep
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import _lib
L = _lib.probe_lib()
grid, iters = 148, 20000
cyc = torch.zeros(grid, dtype=torch.int64, device='cuda')
bgb = torch.zeros(grid, dtype=torch.int64, device='cuda')


def run(n, cg, warps=0, kind=0, gap=0, shift=0, sbo=1024):
    cyc.zero_(); bgb.zero_()
    for rep in range(2):
        _lib.check(L.tdrn_debug_umma_rate_bg(ctypes.c_void_p(cyc.data_ptr()), ctypes.c_void_p(bgb.data_ptr()), grid, n, iters,
                                             cg, warps, kind, gap, shift, sbo), 'rate_bg')
        torch.cuda.synchronize()
    c = cyc[:grid // cg].float().mean().item()
    return c / (iters * 4), bgb.float().mean().item() / max(c, 1.0)


if 'views' not in sys.argv[1:]:
    for n in (64, 128, 256):
        for cg in (1, 2):
            for kind in (0, 1):
                for warps, gap in ((0, 0), (1, 256), (1, 0), (2, 0), (3, 0)):
                    if warps == 0 and kind == 1:
                        continue
                    per, bg = run(n, cg, warps, kind, gap)
                    tc = (4096 + n * 32 // cg) / per                 # operand bytes each SM's tensor core reads per clock
                    print('N=%3d cta_group::%d  background %s x%d warps gap %3d: %6.1f cycles per MMA  tensor-core reads %5.1f B/clk/SM  '
                          'background %5.1f B/clk/SM' % (n, cg, ('loads ', 'stores')[kind], warps, gap, per, tc, bg), flush=True)

# A views of the halo kernels: tap (r, s) of the 8 x 16 tile starts r * 10 + s rows in, 8-row groups 1280 bytes apart
# (conv_halo_kernel); the 16 x 16 super-tile of conv_halo_stream_kernel uses an 18-pixel pitch: shift r * 18 + s, SBO 2304.
for n in (64, 128):
    for cg in (1, 2):
        for shift, sbo in ((0, 1024), (8, 1024), (1, 1024), (11, 1024), (0, 1280), (1, 1280), (10, 1280), (22, 1280)):
            per, _ = run(n, cg, shift=shift, sbo=sbo)
            print('N=%3d cta_group::%d  A start +%2d rows, SBO %4d B: %6.1f cycles per MMA' % (n, cg, shift, sbo, per), flush=True)

# The issue pattern of conv_halo_kernel (9 taps x 4 MMAs per tile, one weight block per tap, halo views of A) without TMA and
# without an epilogue: halo views vs aligned views, with and without the TMEM double-buffer handshake (tdrn_debug_umma_rate_halo).
tiles = 2000
for n in (64, 128):
    for aligned in (0, 1):
        for handshake in (0, 1):
            cyc.zero_()
            for rep in range(2):
                _lib.check(L.tdrn_debug_umma_rate_halo(ctypes.c_void_p(cyc.data_ptr()), grid, n, tiles, aligned, handshake), 'rate_halo')
                torch.cuda.synchronize()
            per = cyc.float().mean().item() / (tiles * 36)
            print('halo pattern N=%3d  %s A views, %s: %6.1f cycles per MMA' %
                  (n, ('halo (shifted, SBO 1280)', 'aligned (SBO 1024)      ')[aligned],
                   ('no handshake        ', 'TMEM double-buffer handshake')[handshake], per), flush=True)
