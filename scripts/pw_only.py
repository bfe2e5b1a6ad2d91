"""ncu / timing target: the pointwise 1x1 convs of the MobileNet-320 b64 trunk alone (half operands, as the trunk runs them)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from tdrn_b200 import ops
g = torch.Generator().manual_seed(0)
B = 64
LAYERS = [(40, 512, 512), (160, 32, 64), (80, 128, 256), (80, 256, 256), (20, 1024, 1024), (40, 256, 512)]
only = [int(a) for a in sys.argv[1:]] if len(sys.argv) > 1 else range(len(LAYERS))
for li in only:
    hw, cin, cout = LAYERS[li]
    x = torch.randn(B, hw, hw, cin, generator=g).to(torch.float16).cuda()
    pc = ops.PackedConv(torch.randn(cout, cin, 1, 1, generator=g) * cin ** -0.5, torch.randn(cout, generator=g), None, 1, 0, 1, device='cuda', want_f16=True)
    run = lambda: ops.conv2d(x, pc, relu=True, use_tc=True)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    mb = B * hw * hw * (cin + cout) * 2 / 1e6
    print('pw %4d->%4d @%3d  %.4f ms  %.0f TFLOP/s  %.2f TB/s (in + out)' % (cin, cout, hw, ms, 2.0 * B * hw * hw * cin * cout / ms / 1e9, mb / ms / 1e3))
