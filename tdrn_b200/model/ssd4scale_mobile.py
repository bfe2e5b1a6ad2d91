"""SSD4Scale_MobNet (TDRN static + temporal nets on the MobileNet trunk): drop-in for the reference's
model/ssd4scale_mobile.py, the pair `evaluate_trn.py:537` builds for ``--version mobile``.

static   (deform=False): plain 3x3 heads (with bias) on the four ARM sources; ``ret_loc`` also returns the raw
                         NCHW loc maps that drive the temporal net's offsets (:113-119,:134-135).
temporal (deform=True) : offsets = offset[k](ref_loc[k]) (1x1, 12 -> 144, dg = 8) or the cached ``offset_list``
                         (:87-93); bias-free ConvOffset2d heads straight on the ARM sources (:108-111).
Trunk and sources are those of the DualRefineDet MobileNet variant (same layer list, :19-51), except that
L2Norm_4_3 starts at scale 10 (:40).
"""
import torch.nn as nn

from ..layers.modules.l2norm import L2Norm
from .dualrefinedet_mobilenet import DW_CFG, mobilenet_sources
from .networks import conv_dw
from .ssd4scale_vgg import SSD4ScaleBase


class SSD4Scale_MobNet(SSD4ScaleBase):
    fp32_tensor_cores = False      # fp32 precision keeps the CUDA-core convs on this trunk (see _engine.Engine.use_x3)
    def __init__(self, size, num_classes=21, phase='train', c7_channel=1024, deform=False):
        super(SSD4Scale_MobNet, self).__init__()
        self.num_classes, self.size, self.phase, self.deform = num_classes, size, phase, deform
        first = nn.Sequential(nn.Conv2d(3, 32, 3, 2, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(inplace=True))
        cfg = DW_CFG[:-1] + [(1024, c7_channel, 1)]
        self.backbone = nn.ModuleList([first] + [conv_dw(i, o, s) for (i, o, s) in cfg])
        self.L2Norm_4_3 = L2Norm(512, 10)
        self.L2Norm_5_3 = L2Norm(1024, 8)
        self.extras = nn.ModuleList([
            nn.Sequential(nn.Conv2d(cin, 256, kernel_size=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                          conv_dw(256, 512, 2)) for cin in (c7_channel, 512)])
        self._add_heads([512, c7_channel, 512, 512], num_classes, deform)

    def _sources(self, E, x):
        return mobilenet_sources(E, x)


def build_net(phase, size=320, num_classes=21, c7_channel=1024, deform=False):
    if size not in [320, 512]:
        print("Error: Sorry only SSD320 and SSD512 is supported currently!")
        return
    return SSD4Scale_MobNet(size, num_classes=num_classes, phase=phase, c7_channel=c7_channel, deform=deform)
