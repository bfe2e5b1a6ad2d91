"""Three eager tdrn_detect calls in regime T (trained-like scores, ~1.2 % candidates) and three in regime R (every prior a
candidate): the target of `ncu --set full -k regex:'detect_front|nms_segment'` (scripts/gpu_profile_detect.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tdrn_b200.layers.functions import Detect, PriorBox
from tdrn_b200.data import mb_cfg
pri = PriorBox(mb_cfg['VOC_320']).forward().cuda()
g = torch.Generator().manual_seed(7)
P, C, B = 6375, 21, 32
loc = torch.randn(B, P, 4, generator=g).cuda(); arm = (0.5 * torch.randn(B, P, 4, generator=g)).cuda()
det = Detect(C, 0, 200, 0.01, 0.45)
for bias, scale in ((7.7, 1.0), (0.0, 0.1)):
    logits = torch.randn(B * P, C, generator=g) * scale; logits[:, 0] += bias
    conf = torch.softmax(logits, 1).cuda()
    for _ in range(3):
        det.forward(loc, conf, pri, arm_loc_data=arm)
    torch.cuda.synchronize()
