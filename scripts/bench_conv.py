"""Micro-benchmark of single conv layers of DualRefineDet-VGGBN-320 b32 through tdrn_conv2d_tc (CUDA events,
20 iterations after 5 warm-ups, inputs 105-420 MB > L2).  Development aid, not a bench.py number."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import ops

B = 32
LAYERS = [  # name, cin, cout, hw, pool
    ('conv1_2', 64, 64, 320, True), ('conv2_1', 64, 128, 160, False), ('conv2_2', 128, 128, 160, True),
    ('conv3_1', 128, 256, 80, False), ('conv3_2', 256, 256, 80, False), ('conv4_1', 256, 512, 40, False),
    ('conv4_2', 512, 512, 40, False), ('conv5_1', 512, 512, 20, False),
    ('tcb0_1', 512, 256, 40, False), ('tcb0_2', 256, 256, 40, False), ('tcb1_1', 512, 256, 20, False), ('tcb2_1', 1024, 256, 10, False),
]
only = sys.argv[1:] if len(sys.argv) > 1 else None
g = torch.Generator().manual_seed(0)
for name, cin, cout, hw, pool in LAYERS:
    if only and name not in only:
        continue
    x = torch.randn(B, hw, hw, cin, generator=g).to(torch.bfloat16).cuda()
    pc = ops.PackedConv(torch.randn(cout, cin, 3, 3, generator=g) * 0.05, torch.randn(cout, generator=g), None, 1, 1, 1, device='cuda')
    run = lambda: ops.conv2d(x, pc, relu=True, use_tc=True, pool=pool)
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
    fl = 2.0 * B * hw * hw * cout * cin * 9
    print('%-8s %3d->%3d @%3d%s  %.4f ms  %.0f TFLOP/s' % (name, cin, cout, hw, ' +pool' if pool else '      ', ms, fl / ms / 1e9))
