"""Multi-scale + horizontal-flip testing with box voting (reference multi_eval.py:496-655, `test_net`), per image:
the image is pre-processed at every scale of `multi_scale[base]`, unflipped and flipped (device resize, tdrn_preprocess),
run through net + Detect with that scale's priors, and the 2 x len(scales) detection sets are merged per class by
tdrn_multiscale_vote (gather + un-flip + size rule + bbox_vote) -- one device->host copy per image instead of the
reference's per-pass, per-class round trips.
"""
import numpy as np
import torch

from .. import ops
from ..data import multi_cfg, multi_cfg_512, multi_scale, preprocess_frames
from ..layers.functions import PriorBox

# (base size, scale) -> (rule, threshold), multi_eval.py:574-625.  rule 0: longer side > thr, 1: shorter side < thr.
# ('320_706' in the reference's table is read as 704, the scale its own list runs, multi_eval.py:21-24 vs :618)
SIZE_RULES = {
    (320, 192): (0, 32), (512, 320): (0, 32), (320, 320): (0, 0), (512, 512): (0, 0),
    (320, 384): (1, 160), (512, 640): (1, 160), (320, 448): (1, 128), (320, 512): (1, 96), (320, 576): (1, 64),
    (320, 704): (1, 32), (512, 1216): (1, 32),
}


class MultiScaleTester(object):
    def __init__(self, net, detector, base_size, mean, scales=None, device='cuda'):
        self.net, self.detector, self.base, self.mean, self.device = net, detector, int(base_size), mean, device
        self.scales = list(scales) if scales is not None else list(multi_scale[str(self.base)])
        cfgs = multi_cfg if self.base == 320 else multi_cfg_512
        self.priors = {}
        for v in self.scales:                                    # multi_eval.py:512-519
            if (self.base, v) not in SIZE_RULES or str(v) not in cfgs:
                raise ValueError('scale %d is not in the multi-scale table of base size %d' % (v, self.base))
            self.priors[v] = PriorBox(cfgs[str(v)]).forward().to(device)

    def passes(self, image):
        """-> (dets [K,C,top_k,5] on the device, flips, rules, thresholds) in the reference's pass order."""
        dets, flips, rules, thrs = [], [], [], []
        with torch.no_grad():
            for v in self.scales:
                for flip in (False, True):                       # multi_eval.py:534-544
                    x = preprocess_frames(image, v, self.mean, to_rgb=True, device=self.device, flip=flip)
                    out = self.net(x)
                    if len(out) == 4:                            # RefineDet family: (arm_loc, _, loc, conf)
                        d = self.detector.forward(out[2], out[3], self.priors[v], arm_loc_data=out[0])
                    else:
                        d = self.detector.forward(out[0], out[1], self.priors[v])
                    dets.append(d[0])
                    flips.append(flip)
                    r, t = SIZE_RULES[(self.base, v)]
                    rules.append(r); thrs.append(t)
        return torch.stack(dets, 0), flips, rules, thrs

    def detect(self, image):
        """image: HWC uint8 (cv2 order).  -> list over classes (index 0 empty) of [m,5] float32 arrays (x1,y1,x2,y2,score) in
        pixels: all_boxes[j][i] of multi_eval.py:641-644."""
        h, w = int(image.shape[0]), int(image.shape[1])
        dets, flips, rules, thrs = self.passes(image)
        rows, cnt = ops.multiscale_vote(dets, flips, rules, thrs, w, h)
        rows, cnt = rows.cpu().numpy(), cnt.cpu().numpy()
        return [rows[j, :cnt[j]].copy() for j in range(rows.shape[0])]
