"""Micro-benchmark of tdrn_detect on the bench workload's own loc/conf tensors (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tdrn_b200 import ops
from tdrn_b200.layers.functions import Detect, PriorBox
from tdrn_b200.data import mb_cfg
from tdrn_b200.utils.synthetic import frames

net = bench.build_synthetic_net().cuda().set_precision('bf16')
x = frames(32, 320, 100).cuda()
with torch.no_grad():
    arm_loc, _, loc, conf = net(x)
pri = PriorBox(mb_cfg['VOC_320']).forward().cuda()
print('candidates/(img,class): mean %.0f' % ((conf.view(32, -1, 21)[:, :, 1:] > 0.01).float().sum(1).mean().item()))

def t(fn, n=20):
    """CUDA-graph replay timing: Detect is a memset + three short kernels, eager launches would measure the host."""
    return bench.graph_replay_ms(fn, torch.cuda.Stream(), iters=n)

print('decode only            %.4f ms' % t(lambda: ops.decode(loc, pri, arm_loc)))
for name, kw in (('full (0.01,0.45,200)', dict(top_k=200, conf_thresh=0.01, nms_thresh=0.45)),
                 ('no candidates (0.99)', dict(top_k=200, conf_thresh=0.99, nms_thresh=0.45)),
                 ('top_k=1 (sort only)', dict(top_k=1, conf_thresh=0.01, nms_thresh=0.45)),
                 ('conf 0.05', dict(top_k=200, conf_thresh=0.05, nms_thresh=0.45)),
                 ('conf 0.2', dict(top_k=200, conf_thresh=0.2, nms_thresh=0.45))):
    d = Detect(21, 0, kw['top_k'], kw['conf_thresh'], kw['nms_thresh'])
    ms = t(lambda: d.forward(loc, conf, pri, arm_loc_data=arm_loc))
    out = d.forward(loc, conf, pri, arm_loc_data=arm_loc)
    print('%-24s %.4f ms   kept/(img,class) %.1f' % (name, ms, (out[..., 0] > 0).float().sum(-1)[:, 1:].mean().item()))

# regime T: trained-like scores, background logit +7.7
g = torch.Generator().manual_seed(7)
logits = torch.randn(32 * 6375, 21, generator=g); logits[:, 0] += 7.7
conf_t = torch.softmax(logits, 1).cuda()
loc_t = torch.randn(32, 6375, 4, generator=g).cuda(); arm_t = (0.5 * torch.randn(32, 6375, 4, generator=g)).cuda()
d = Detect(21, 0, 200, 0.01, 0.45)
ms = t(lambda: d.forward(loc_t, conf_t, pri, arm_loc_data=arm_t))
nb = 32 * (6375 * (32 + 84) + 21 * 200 * 20) + 6375 * 16
print('regime T (%.2f %% candidates, max %d per segment)  %.4f ms = %.0f GB/s algorithmic' % (
    100 * (conf_t[:, 1:] > 0.01).float().mean().item(), int((conf_t.view(32, -1, 21)[:, :, 1:] > 0.01).sum(1).max()), ms, nb / ms / 1e6))
