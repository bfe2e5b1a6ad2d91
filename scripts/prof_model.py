"""Per-operator CUDA-event breakdown of one forward (+Detect) of a named model (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdrn_b200 import ops
from tdrn_b200.data import mb_cfg
from tdrn_b200.layers.functions import Detect, PriorBox
from tdrn_b200.utils.synthetic import frames, randomize_
name, B = sys.argv[1], int(sys.argv[2])
dev = torch.device('cuda')
if name == 'mobilenet':
    from tdrn_b200.model import dualrefinedet_mobilenet as M
    net = randomize_(M.build_net('test', 320, num_classes=21, def_groups=1, multihead=False), 0).eval().to(dev)
    C, size, cfg, topk = 21, 320, 'VOC_320', 200
elif name == 'coco512':
    from tdrn_b200.model import dualrefinedet_vggbn as V
    net = randomize_(V.build_net('test', 512, num_classes=81, def_groups=1, bn=True, multihead=True), 0).eval().to(dev)
    C, size, cfg, topk = 81, 512, 'VOC_512_RefineDet', 100
elif name == 'tdrn':
    from tdrn_b200.model import ssd4scale_vgg as S
    stat = randomize_(S.build_net('test', 320, num_classes=31, bn=True, deform=False), 0).eval().to(dev)
    net = randomize_(S.build_net('test', 320, num_classes=31, bn=True, deform=True), 1).eval().to(dev)
    stat.engine().multi_stream = False
    C, size, cfg, topk = 31, 320, 'VOC_320', 200
net.engine().multi_stream = False
pri = PriorBox(mb_cfg[cfg]).forward().to(dev)
det = Detect(C, 0, topk, 0.01, 0.45)
x = frames(B, size, 3).to(dev)


def step():
    if name == 'tdrn':
        keys = x[::4]
        s_loc, s_conf, maps = stat(keys, ret_loc=True)
        out = net(x, ref_loc=[m.repeat_interleave(4, 0) for m in maps], ret_off=True)
        return det.forward(out[0], out[1], pri, arm_loc_data=s_loc.repeat_interleave(4, 0))
    a, _, l, c = net(x)
    return det.forward(l, c, pri, arm_loc_data=a, scale=[float(size)] * 4)


with torch.no_grad():
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ops.prof_begin()
    step()
    rec = ops.prof_end()
tot = sum(r[2] for r in rec)
print('total timed ops %.3f ms (%d ops)' % (tot, len(rec)))
for lab, work, ms in sorted(rec, key=lambda r: -r[2])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print('%-50s %8.4f ms  %8.1f G/s' % (lab, ms, work / ms / 1e6))
