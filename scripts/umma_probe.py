"""Decode the shared-memory addressing of SWIZZLE_128B K-major UMMA descriptors with unaligned start rows and
arbitrary SBO (tdrn_debug_umma_probe).  Prints, per variant, whether A[m][k] came from row r0 + (m/8)*(sbo/128) + m%8,
column k, and if not, the first few mismatches."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import _lib

L = _lib.probe_lib()
out = torch.empty(128, 64, device='cuda')
rows = 400


def run(r0, sbo, bo, mode):
    _lib.check(L.tdrn_debug_umma_probe(ctypes.c_void_p(out.data_ptr()), r0, sbo, bo, mode, rows), 'probe')
    return out.cpu().clone()


for r0 in (0, 1, 2, 3, 10, 11, 21):
    for sbo in (1024, 1280, 2048):
        for bo_mode in ('zero', 'doc'):
            bo = 0 if bo_mode == 'zero' else (r0 & 7)
            R = run(r0, sbo, bo, 0)
            K = run(r0, sbo, bo, 1)
            m = torch.arange(128)
            exp_row = (r0 + (m // 8) * (sbo // 128) + m % 8).float().view(128, 1).expand(128, 64)
            exp_col = torch.arange(64).float().view(1, 64).expand(128, 64)
            ok_r = torch.equal(R, exp_row)
            ok_k = torch.equal(K, exp_col)
            msg = 'r0=%2d sbo=%4d base_offset=%d(%s): rows %s cols %s' % (r0, sbo, bo, bo_mode, 'OK ' if ok_r else 'BAD', 'OK ' if ok_k else 'BAD')
            if not (ok_r and ok_k):
                bad = ((R != exp_row) | (K != exp_col)).nonzero()[:4].tolist()
                msg += '  e.g. ' + ', '.join('(m=%d,k=%d)->row %g col %g' % (a, b, R[a, b], K[a, b]) for a, b in bad)
            print(msg)
