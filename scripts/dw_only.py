"""ncu target: the depthwise 3x3 kernels alone on the MobileNet-320 b64 shapes, in the trunk's format (IEEE half: the packed-half
kernels; pass bf16 as the first argument for the fp32-accumulating bf16 kernel).  No timing, two launches per shape."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from tdrn_b200 import ops
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == 'bf16' else torch.float16
bn = lambda c: [torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]
for (B, H, W, C, s) in [(64, 40, 40, 512, 1), (64, 80, 80, 256, 1), (64, 160, 160, 64, 2)]:
    x = torch.randn(B, H, W, C, generator=g).to(dt).cuda()
    pd = ops.PackedDw(torch.randn(C, 1, 3, 3, generator=g) * 0.3, bn(C), s, 'cuda')
    for _ in range(2):
        ops.dwconv3x3(x, pd, relu=True)
torch.cuda.synchronize()
