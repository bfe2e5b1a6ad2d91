"""Deterministic synthetic weights / frames for benchmarks (no datasets or checkpoints offline).

He-uniform convolutions, xavier-uniform deformable heads (model/networks.py:727), *randomised*
BatchNorm statistics so that BN folding is exercised, damped conv7 so the un-normalised ARM sources
stay O(1) (keeps ARM regressions / learned offsets at a few pixels, like a trained net).
"""
import math

import torch
import torch.nn as nn

from ..model.networks import ConvOffset2d


def randomize_(module, seed=0):
    g = torch.Generator().manual_seed(seed)

    def uni(t, bound):
        t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * bound)

    with torch.no_grad():
        for name, m in module.named_modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
                gain = 1.0
                if name.startswith('offset'):
                    gain = 1.5
                elif isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1) and name.startswith('backbone') and m.out_channels >= 512:
                    gain = 0.04
                uni(m.weight, gain * math.sqrt(6.0 / fan_in))
                if m.bias is not None:
                    uni(m.bias, 0.1)
            elif isinstance(m, ConvOffset2d):
                fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
                fan_out = m.weight.shape[0] * m.weight.shape[2] * m.weight.shape[3]
                uni(m.weight, math.sqrt(6.0 / (fan_in + fan_out)))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return module


def frames(batch, size, seed):
    """[B,3,S,S] fp32 N(0,1) (mean-subtracted pixel statistics are irrelevant to cost)."""
    return torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(seed))
