"""CPU restatement of PriorBox / decode / center_size / Detect.  TEST INFRASTRUCTURE ONLY.

Follows:
  * PriorBox.forward          layers/functions/prior_box.py:33-64
  * decode                    layers/box_utils.py:176-195
  * center_size               layers/box_utils.py:16-25
  * Detect.__init__/forward   layers/functions/detection.py:14-23,25-70

Pinned (tests/test_oracle_vs_reference.py, tests/golden/*) against the reference's own Python
files executed in the build container through oracle/ref_shim.py.
"""
from itertools import product
from math import sqrt

import numpy as np
import torch

from .nms_ref import cpu_nms

VARIANCE = (0.1, 0.2)                                      # detection.py:23

# Values of the two prior dictionaries the hot path uses (data/config.py:57-68, :70-81).
VOC_320 = {
    'feature_maps': [40, 20, 10, 5], 'min_dim': 320, 'steps': [8, 16, 32, 64],
    'min_sizes': [32, 64, 128, 256], 'max_sizes': [], 'aspect_ratios': [[2], [2], [2], [2]],
    'variance': [0.1, 0.2], 'clip': True, 'flip': True, 'name': 'VOC_320',
}
VOC_512_RefineDet = {
    'feature_maps': [64, 32, 16, 8], 'min_dim': 512, 'steps': [8, 16, 32, 64],
    'min_sizes': [32, 64, 128, 256], 'max_sizes': [], 'aspect_ratios': [[2], [2], [2], [2]],
    'variance': [0.1, 0.2], 'clip': True, 'flip': True, 'name': 'VOC_512_RefineDet',
}


def prior_box(cfg):
    """prior_box.py:33-64 -- float64 Python arithmetic, then one fp32 conversion, then clamp."""
    for v in (cfg['variance'] or [0.1]):
        if v <= 0:
            raise ValueError('Variances must be greater than 0')    # prior_box.py:29-31
    image_size = cfg['min_dim']
    mean = []
    for k, f in enumerate(cfg['feature_maps']):
        for i, j in product(range(f), repeat=2):
            f_k = image_size / cfg['steps'][k]
            cx = (j + 0.5) / f_k
            cy = (i + 0.5) / f_k
            s_k = cfg['min_sizes'][k] / image_size
            mean += [cx, cy, s_k, s_k]
            if len(cfg['max_sizes']):
                s_k_prime = sqrt(s_k * (cfg['max_sizes'][k] / image_size))
                mean += [cx, cy, s_k_prime, s_k_prime]
            for ar in cfg['aspect_ratios'][k]:
                mean += [cx, cy, s_k * sqrt(ar), s_k / sqrt(ar)]
                if cfg['flip']:
                    mean += [cx, cy, s_k / sqrt(ar), s_k * sqrt(ar)]
    out = torch.tensor(mean, dtype=torch.float64).to(torch.float32).view(-1, 4)
    if cfg['clip']:
        out.clamp_(max=1, min=0)
    return out


def decode(loc, priors, variances=VARIANCE):
    """box_utils.py:190-195 (same operation order, fp32)."""
    cxcy = priors[:, :2] + loc[:, :2] * variances[0] * priors[:, 2:]
    wh = priors[:, 2:] * torch.exp(loc[:, 2:] * variances[1])
    x1y1 = cxcy - wh / 2
    x2y2 = wh + x1y1
    return torch.cat((x1y1, x2y2), 1)


def center_size(boxes):
    """box_utils.py:24-25."""
    return torch.cat(((boxes[:, 2:] + boxes[:, :2]) / 2, boxes[:, 2:] - boxes[:, :2]), 1)


def decode_two_stage(loc, priors, arm_loc=None):
    """detection.py:43-48 for one image."""
    default = center_size(decode(arm_loc, priors)) if arm_loc is not None else priors
    return decode(loc, default)


def detect_from_boxes(decoded, conf, scale, num_classes, top_k, conf_thresh, nms_thresh):
    """detection.py:37-63 given already-decoded boxes [B,P,4] and conf [B*P,C] (fp32, CPU)."""
    num, num_priors = decoded.shape[0], decoded.shape[1]
    output = torch.zeros(num, num_classes, top_k, 5)
    conf_preds = conf.view(num, num_priors, num_classes).transpose(2, 1)
    scale_np = scale.numpy().astype(np.float32)
    for i in range(num):
        boxes_i = decoded[i]
        for cl in range(1, num_classes):
            sc = conf_preds[i, cl]
            c_mask = sc.gt(conf_thresh)                             # strict, fp32 compare (:53)
            scores = sc[c_mask]
            if scores.size(0) == 0:
                continue
            boxes = boxes_i[c_mask]
            c_dets = np.hstack((boxes.numpy() * scale_np, scores.numpy()[:, None])).astype(np.float32)
            keep = cpu_nms(c_dets, nms_thresh, max_keep=top_k)      # :60 (early exit == keep[:top_k])
            n = min(len(keep), top_k)
            k = torch.as_tensor(keep[:n], dtype=torch.int64)
            output[i, cl, :n] = torch.cat((scores[k].unsqueeze(1), boxes[k]), 1)   # :61-63
    # detection.py:65-68 is a no-op (fills a copy produced by advanced indexing).
    return output


def detect(loc, conf, priors, arm_loc=None, scale=None, num_classes=21, top_k=200,
           conf_thresh=0.01, nms_thresh=0.45):
    if nms_thresh <= 0:
        raise ValueError('nms_threshold must be non negative.')     # detection.py:20-21
    if scale is None:
        scale = torch.tensor([320, 320, 320, 320], dtype=torch.float32)   # detection.py:25
    num = loc.shape[0]
    decoded = torch.stack([decode_two_stage(loc[i], priors, None if arm_loc is None else arm_loc[i])
                           for i in range(num)])
    return detect_from_boxes(decoded, conf, scale, num_classes, top_k, conf_thresh, nms_thresh)
