import os, sys
sys.path.insert(0, '/root/repo')
import torch
from tdrn_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(32, 3, 320, 320, generator=g).cuda()
pc1 = ops.PackedConv(torch.randn(64, 3, 3, 3, generator=g) * 0.2, torch.randn(64, generator=g) * 0.1, None, 1, 1, 1, device='cuda')
pc2 = ops.PackedConv(torch.randn(64, 64, 3, 3, generator=g) * 0.05, torch.randn(64, generator=g) * 0.1, None, 1, 1, 1, device='cuda')
for _ in range(3):
    y = ops.conv_stem_pair(x, pc1, pc2)
torch.cuda.synchronize()
