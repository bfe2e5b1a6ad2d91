"""Detect mirror (layers/functions/detection.py:8-70) over the C ABI (tdrn_detect)."""
import torch

from ... import ops


class Detect(object):
    """At test time, Detect is the final layer of SSD: two-stage decode, per-class score threshold,
    NMS, top-k.  Same constructor and forward signature as the reference.

    Deviations (documented in INTEGRATION.md): the result lives on the inputs' CUDA device (the
    reference builds it with torch.zeros on the host), and equal scores are ordered "lower prior
    index first" (the reference's order among ties is whatever NumPy's argsort yields).
    """

    def __init__(self, num_classes, bkg_label, top_k, conf_thresh, nms_thresh):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.top_k = top_k
        self.nms_thresh = nms_thresh
        if nms_thresh <= 0:
            raise ValueError('nms_threshold must be non negative.')
        self.conf_thresh = conf_thresh
        self.variance = [0.1, 0.2]

    def forward(self, loc_data, conf_data, prior_data, arm_loc_data=None, scale=None):
        if scale is None:
            scale = [320.0, 320.0, 320.0, 320.0]                    # detection.py:25 default
        elif torch.is_tensor(scale):
            scale = scale.detach().float().cpu().tolist()
        dev = loc_data.device
        if not loc_data.is_cuda:
            raise NotImplementedError('Detect needs CUDA tensors: tdrn_b200 has no CPU path')
        prior_data = prior_data.to(dev)
        if prior_data.dim() == 3:
            prior_data = prior_data[0]
        num = loc_data.size(0)
        num_priors = prior_data.size(0)
        conf = conf_data.reshape(num * num_priors, self.num_classes)
        return ops.detect(loc_data.reshape(num, num_priors, 4), conf, prior_data,
                          None if arm_loc_data is None else arm_loc_data.reshape(num, num_priors, 4),
                          scale, self.num_classes, self.top_k, self.conf_thresh, self.nms_thresh)

    __call__ = forward
