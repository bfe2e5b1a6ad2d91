"""Layer factory mirror of the reference's model/networks.py (hot subset only).

  vgg()            model/networks.py:136-163   (parameter containers; compute happens in _engine)
  conv_dw()        model/networks.py:736-745
  ConvOffset2d     model/networks.py:699-733   -> tdrn_deform_conv_forward (C ABI)
  conv_offset2d    model/networks.py:600-615
"""
import torch
import torch.nn as nn
import torch.nn.init as init
from torch.nn.modules.utils import _pair

from .. import ops

vgg_base = {
    '320': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512],
    '512': [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512],
}


def vgg(cfg, i, batch_norm=False, pool5_ds=False, c7_channel=1024):
    """Same module list (hence the same state-dict indices) as the reference's vgg()."""
    mods, cin = [], i

    def block(conv):
        out = [conv]
        if batch_norm:
            out.append(nn.BatchNorm2d(conv.out_channels))
        out.append(nn.ReLU(inplace=True))
        return out

    for v in cfg:
        if v in ('M', 'C'):
            mods.append(nn.MaxPool2d(kernel_size=2, stride=2, ceil_mode=(v == 'C')))
        else:
            mods += block(nn.Conv2d(cin, v, kernel_size=3, padding=1))
            cin = v
    mods.append(nn.MaxPool2d(kernel_size=2, stride=2, padding=0) if pool5_ds
                else nn.MaxPool2d(kernel_size=3, stride=1, padding=1))
    mods += block(nn.Conv2d(512, 1024, kernel_size=3, padding=6, dilation=6))
    mods += block(nn.Conv2d(1024, c7_channel, kernel_size=1))
    return mods


def conv_dw(inp, oup, stride):
    return nn.Sequential(
        nn.Conv2d(inp, inp, kernel_size=3, stride=stride, padding=1, groups=inp, bias=False),
        nn.BatchNorm2d(inp), nn.ReLU(inplace=True),
        nn.Conv2d(inp, oup, 1, 1, 0, bias=False),
        nn.BatchNorm2d(oup), nn.ReLU(inplace=True))


def conv_offset2d(input, offset, weight, stride=1, padding=0, dilation=1, deform_groups=1):
    if input is not None and input.dim() != 4:
        raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
    return ops.deform_conv_nchw(input, offset, weight, _pair(stride), _pair(padding), _pair(dilation),
                                deform_groups)


class ConvOffset2d(nn.Module):
    """Deformable conv v1 (no bias).  forward(input NCHW fp32 cuda, offset [B, dg*2*kh*kw, Ho, Wo])."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 num_deformable_groups=1):
        super(ConvOffset2d, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.num_deformable_groups = num_deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.weight.data)          # networks.py:727

    def forward(self, input, offset):
        return conv_offset2d(input, offset, self.weight, self.stride, self.padding, self.dilation,
                             self.num_deformable_groups)
