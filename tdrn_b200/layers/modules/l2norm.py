"""L2Norm mirror (layers/modules/l2norm.py:5-21): parameter container + standalone NCHW forward."""
import torch
import torch.nn as nn
import torch.nn.init as init

from ... import ops


class L2Norm(nn.Module):
    def __init__(self, n_channels, scale):
        super(L2Norm, self).__init__()
        self.n_channels = n_channels
        self.gamma = scale or None
        self.eps = 1e-10
        self.weight = nn.Parameter(torch.empty(self.n_channels))
        self.reset_parameters()

    def reset_parameters(self):
        init.constant_(self.weight, self.gamma)

    def forward(self, x):
        y = ops.l2norm(ops.nchw_f32_to_nhwc(x.float(), torch.float32), self.weight.detach().float().contiguous())
        return ops.nhwc_to_nchw_f32(y)
