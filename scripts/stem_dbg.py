import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdrn_b200 import ops
g = torch.Generator().manual_seed(1)
x = torch.randn(2, 3, 8, 64, generator=g)
wt = torch.randn(64, 3, 3, 3, generator=g) * 0.3
bias = torch.randn(64, generator=g)
pc = ops.PackedConv(wt, bias, None, 1, 1, 1, device='cuda', want_bf16=False)
out = ops.conv_first(x.cuda(), pc, True, torch.bfloat16)
torch.cuda.synchronize()
print('ok', out.float().abs().mean().item())
