import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdrn_b200 import _lib
L = _lib.probe_lib()
B, H, W = 2, 8, 64
pw, ph, cx, cy, b, rank, promo = [int(v) for v in sys.argv[1:8]]
x = torch.arange(B * 3 * H * W, dtype=torch.float32).view(B, 3, H, W).cuda()
out = torch.full((pw * ph * 3,), -7.0, device='cuda')
rc = L.tdrn_debug_tma_f32(ctypes.c_void_p(x.data_ptr()), B, H, W, pw, ph, cx, cy, b, rank, promo, ctypes.c_void_p(out.data_ptr()))
if rc:
    print('FAIL', L.tdrn_last_error().decode()[:120]); sys.exit(0)
o = out.cpu().view(3, ph, pw)
xp = torch.zeros(3, H + 40, W + 80); xp[:, 20:20 + H, 20:20 + W] = x[b].cpu()
exp = xp[:, 20 + cy:20 + cy + ph, 20 + cx:20 + cx + pw]
print('OK match=%s' % torch.equal(o, exp))
