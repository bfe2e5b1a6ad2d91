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

def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

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
