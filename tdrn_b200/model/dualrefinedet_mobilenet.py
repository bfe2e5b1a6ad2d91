"""DualRefineDet-MobileNet: drop-in for the reference's model/dualrefinedet_mobilenet.py.

``build_net(phase, size, num_classes, def_groups, multihead)`` (:210-214); state-dict keys :19-121;
forward outputs (:183-189): (arm_loc, None, odm_loc, softmax(conf)).
"""
import os

import torch
import torch.nn as nn

from ..layers.modules.l2norm import L2Norm
from ._base import DetectorBase
from ._engine import prior_layout
from .dualrefinedet_vggbn import add_fpn, add_deform_heads, _list
from .networks import conv_dw
from .. import ops

DW_CFG = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 1), (256, 256, 1), (256, 512, 2),
          (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1),
          (512, 1024, 2), (1024, 1024, 1)]


def _dw_block(E, name, x, stride, out_dtype=None):
    """conv_dw (networks.py:736-745): depthwise 3x3+BN+ReLU then pointwise 1x1+BN+ReLU.  ``out_dtype``: format of the block's
    output when it differs from the input's (the half-precision trunk hands its last source to the bf16 heads)."""
    pd = E.packed_dw(name + '.0', name + '.1', stride)
    if E.use_tc and x.dtype == torch.bfloat16 and stride == 1 and os.environ.get('TDRN_DWPW', '0') == '1':
        # OPT-IN (TDRN_DWPW=1): both halves in one kernel, the depthwise output stays in shared memory (tdrn_conv_dwpw; same
        # bits as the two kernels).  Measured on B200 (b64, scripts/dwpw_timing.py): slower than the two kernels on every layer
        # (512 -> 512 @40x40: 0.249 ms against 0.081 + 0.074) -- the depthwise arithmetic (bf16 unpack + fp32 FMA, ~2 300 issue cycles
        # per 64-channel block on 8 producer warps) is 4-5x the block's MMA time, so the fused CTA runs at CUDA-core speed with
        # 10 of its 19 warps waiting, while the stand-alone depthwise kernel fills all four schedulers of every SM.  The accuracy
        # argument for fusing does not hold either (scripts/mobilenet_bf16_emulation.py: the MMA operand is bf16 in both forms).
        pc = E.packed(name + '.3', 1, 0, 1, name + '.4')
        if pc.w_bf16 is not None and x.shape[3] % 8 == 0 and pc.cout % 8 == 0:
            return ops.conv_dwpw(x, pd, pc)
    x = ops.dwconv3x3(x, pd, relu=True)
    if out_dtype is not None and out_dtype != x.dtype:
        return E.conv(name + '.3', x, bn=name + '.4', relu=True, out_dtype=out_dtype)
    return E.conv(name + '.3', x, bn=name + '.4', relu=True)


def mobilenet_sources(E, x):
    """The four ARM sources of the MobileNet trunk (dualrefinedet_mobilenet.py:139-152; the same loop in
    ssd4scale_mobile.py:99-110): L2Norm of backbone[11]'s output (512 @ 40x40), L2Norm of the last block
    (1024 @ 20x20), then the two 1x1 + conv_dw(s2) extras (512 @ 10x10, 512 @ 5x5).  NHWC."""
    # 16-bit mode: the trunk (stem, 13 conv_dw blocks, the first extras block) computes and stores IEEE half -- activations AND
    # weights -- and hands its four sources to the bf16 ARM heads / TCB as bf16: the two L2Norms convert, the first extras block's
    # pointwise conv writes bf16, the second extras block is bf16 throughout.  (CPU emulation of exactly this plan: conf 1.3e-2
    # against 3.0e-2 with bf16 storage, scripts/mobilenet_bf16_emulation.py.)  TDRN_MOBILE_BF16=1 keeps the bf16 trunk.
    half = E.half_trunk and E.act == torch.bfloat16 and ops.conv_first_f16_ok(x)
    if half:
        x = E.conv_first('backbone.0.0', x, 2, 'backbone.0.1', out_dtype=torch.float16)
    else:
        x = E.conv_first('backbone.0.0', x, 2, 'backbone.0.1')
    arm_sources = []
    for n, (i, o, s) in enumerate(DW_CFG):
        if n + 1 == 12:
            arm_sources.append(ops.l2norm(x, E.vec('L2Norm_4_3.weight'), out_dtype=E.act))
        x = _dw_block(E, 'backbone.%d' % (n + 1), x, s)
    arm_sources.append(ops.l2norm(x, E.vec('L2Norm_5_3.weight'), out_dtype=E.act))
    for e in range(2):
        x = E.conv('extras.%d.0' % e, x, bn='extras.%d.1' % e, relu=True)
        x = _dw_block(E, 'extras.%d.3' % e, x, 2, out_dtype=E.act)
        arm_sources.append(x)
    return arm_sources


class RefineSSD(DetectorBase):
    fp32_tensor_cores = False      # fp32 precision keeps the CUDA-core convs on this trunk (see _engine.Engine.use_x3)
    def __init__(self, size, num_classes=21, phase='train', def_groups=1, multihead=False):
        super(RefineSSD, self).__init__()
        self.num_classes, self.size, self.phase = num_classes, size, phase
        self.def_groups, self.multihead = def_groups, multihead
        first = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True))
        self.backbone = nn.ModuleList([first] + [conv_dw(i, o, s) for (i, o, s) in DW_CFG])
        self.L2Norm_4_3 = L2Norm(512, 20)
        self.L2Norm_5_3 = L2Norm(1024, 8)
        self.extras = nn.ModuleList([
            nn.Sequential(nn.Conv2d(cin, 256, kernel_size=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                          conv_dw(256, 512, 2)) for cin in (1024, 512)])
        src = [512, 1024, 512, 512]
        add_fpn(self, src, bias=False)
        self.arm_loc = _list(lambda k: nn.Conv2d(src[k], 12, kernel_size=3, stride=1, padding=1, bias=False))
        add_deform_heads(self, num_classes, def_groups, multihead, False)
        if phase == 'test':
            self.softmax = nn.Softmax(dim=1)

    def forward(self, x, _offsets=None):
        """``_offsets``: test hook, see dualrefinedet_vggbn.RefineSSD.forward."""
        E = self.engine()
        arm_sources = mobilenet_sources(E, self._check_input(x))
        P, lv = prior_layout(arm_sources)
        arm_loc, offs, offs2, odm_sources = E.arm_and_tcb(arm_sources, P, lv, self.multihead)
        if _offsets is not None:
            offs = [ops.nchw_f32_to_nhwc(o.float(), torch.float32) for o in _offsets[0]]
            offs2 = [ops.nchw_f32_to_nhwc(o.float(), torch.float32) for o in (_offsets[1] or [])]
        odm_loc, conf = E.deform_heads(odm_sources, offs, offs2, P, lv, self.num_classes, self.def_groups,
                                       self.multihead)
        return arm_loc, None, odm_loc, conf


def build_net(phase, size=320, num_classes=21, def_groups=1, multihead=False):
    if size not in [320, 512]:
        print("Error: Sorry only SSD320 and SSD512 is supported currently!")
        return
    return RefineSSD(size, num_classes=num_classes, phase=phase, def_groups=def_groups, multihead=multihead)
