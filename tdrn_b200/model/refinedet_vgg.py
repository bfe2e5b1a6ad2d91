"""RefineDet-VGG16 (plain-conv ODM heads): drop-in for the reference's model/refinedet_vgg.py
(BASELINE config 1).  ``build_net`` :230-235; outputs :198-210.
"""
import torch.nn as nn

from ..layers.modules.l2norm import L2Norm
from ._base import DetectorBase
from ._engine import prior_layout
from .dualrefinedet_vggbn import add_fpn, add_vgg_extras, _list
from .networks import vgg, vgg_base


class RefineSSD(DetectorBase):
    def __init__(self, size, num_classes=21, use_refine=False, phase='train', c7_channel=1024, bn=False,
                 multihead=False):
        super(RefineSSD, self).__init__()
        self.num_classes, self.size, self.use_refine, self.phase = num_classes, size, use_refine, phase
        self.bn, self.multihead = bn, multihead
        self.backbone = nn.ModuleList(vgg(vgg_base['320'], 3, batch_norm=bn, pool5_ds=True, c7_channel=c7_channel))
        self.L2Norm_4_3 = L2Norm(512, 10)
        self.L2Norm_5_3 = L2Norm(512, 8)
        src = [512, 512, c7_channel, 512]
        add_fpn(self, src)
        add_vgg_extras(self, bn, c7_channel)
        c = lambda i, o, k: nn.Conv2d(i, o, kernel_size=k, stride=1, padding=k // 2)
        if use_refine:
            self.arm_loc = _list(lambda k: c(src[k], 12, 3))
        self.odm_loc = _list(lambda k: c(256, 12, 3))
        self.odm_conf = _list(lambda k: c(256, 3 * num_classes, 3))
        if multihead:
            self.odm_loc_2 = _list(lambda k: c(256, 12, 5))
            self.odm_conf_2 = _list(lambda k: c(256, 3 * num_classes, 5))
        if phase == 'test':
            self.softmax = nn.Softmax(dim=1)

    def forward(self, x):
        E = self.engine()
        x = self._check_input(x)
        arm_sources = E.vgg_trunk(x, self.bn)
        P, lv = prior_layout(arm_sources)
        arm_loc = None
        if self.use_refine:
            arm_loc, _, _ = E.arm_heads(arm_sources, P, lv, False, with_offsets=False)
        odm_sources = E.fpn(arm_sources)
        odm_loc, conf = E.plain_heads(odm_sources, P, lv, self.num_classes, self.multihead, 'odm_loc', 'odm_conf')
        if self.use_refine:
            return arm_loc, None, odm_loc, conf
        return odm_loc, conf


def build_net(phase, size=320, num_classes=21, use_refine=False, c7_channel=1024, bn=False, multihead=False):
    if size not in [320, 512]:
        print("Error: Sorry only SSD300 and SSD512 is supported currently!")
        return
    return RefineSSD(size, num_classes=num_classes, use_refine=use_refine, phase=phase, c7_channel=c7_channel,
                     bn=bn, multihead=multihead)
