"""SSD4Scale (TDRN static + temporal nets): drop-in for the reference's model/ssd4scale_vgg.py.

static   (deform=False): plain 3x3 heads on the ARM sources; ``ret_loc`` also returns the raw NCHW
                         loc maps that drive the temporal net's offsets (:111-115,:130-131).
temporal (deform=True) : offsets = offset[k](ref_loc[k]) (1x1, 12 -> 144, dg = 8) or the cached
                         ``offset_list``; ConvOffset2d heads straight on the ARM sources (:72-78,:106-109).
"""
import torch
import torch.nn as nn

from .. import ops
from ..layers.modules.l2norm import L2Norm
from ._base import DetectorBase
from ._engine import prior_layout
from .dualrefinedet_vggbn import add_vgg_extras, _list
from .networks import vgg, vgg_base, ConvOffset2d

DF_GROUP = 8


class SSD4ScaleBase(DetectorBase):
    """Head construction and ``forward`` shared by the VGG and MobileNet SSD4Scale nets; a subclass builds its
    trunk and says how the four ARM sources come out of it (``_sources``)."""

    def _add_heads(self, src, num_classes, deform):
        if deform:
            self.offset = _list(lambda k: nn.Conv2d(12, DF_GROUP * 18, kernel_size=1))
            self.arm_loc = _list(lambda k: ConvOffset2d(src[k], 12, 3, 1, 1, num_deformable_groups=DF_GROUP))
            self.arm_conf = _list(lambda k: ConvOffset2d(src[k], 3 * num_classes, 3, 1, 1, num_deformable_groups=DF_GROUP))
        else:
            self.arm_loc = _list(lambda k: nn.Conv2d(src[k], 12, kernel_size=3, stride=1, padding=1))
            self.arm_conf = _list(lambda k: nn.Conv2d(src[k], 3 * num_classes, kernel_size=3, stride=1, padding=1))
        if self.phase == 'test':
            self.softmax = nn.Softmax(dim=1)

    def _sources(self, E, x):
        raise NotImplementedError

    def trunk(self, x):
        """The four ARM sources of ``x`` alone (not part of the reference's surface): lets a streaming loop run the temporal net's
        trunk next to the static net of the same key frame and hand the sources to ``forward(..., _sources=...)``."""
        return self._sources(self.engine(), self._check_input(x))

    def forward(self, x, ref_loc=list(), offset_list=list(), ret_loc=False, ret_off=False, _sources=None):
        E = self.engine()
        x = self._check_input(x)
        offs_nhwc = None
        if self.deform:
            if not offset_list:
                offs_nhwc = [E.conv('offset.%d' % k, ops.nchw_f32_to_nhwc(rl.float(), torch.float32),
                                    out_dtype=torch.float32) for k, rl in enumerate(ref_loc)]
            else:
                offs_nhwc = [ops.nchw_f32_to_nhwc(o.float(), torch.float32) for o in offset_list]
        src = _sources if _sources is not None else self._sources(E, x)
        P, lv = prior_layout(src)
        if self.deform:
            loc, conf = E.deform_heads(src, offs_nhwc, None, P, lv, self.num_classes, DF_GROUP, False,
                                       loc_name='arm_loc', conf_name='arm_conf')
        else:
            loc, conf = E.plain_heads(src, P, lv, self.num_classes, False, 'arm_loc', 'arm_conf')
        out = [loc, conf]
        if ret_loc:                         # raw NCHW loc maps: un-flatten the NHWC rows of each level
            B = loc.shape[0]
            maps = []
            for k, s in enumerate(src):
                H, W = s.shape[1], s.shape[2]
                lvl = loc.view(B, P * 4)[:, lv[k] * 4:(lv[k] + H * W * 3) * 4].reshape(B, H, W, 12)
                maps.append(ops.nhwc_to_nchw_f32(lvl.contiguous()))
            out.append(maps)
        if ret_off:
            out.append([ops.nhwc_to_nchw_f32(o) for o in offs_nhwc])
        return tuple(out)


class SSD4Scale(SSD4ScaleBase):
    def __init__(self, size, num_classes=21, phase='train', c7_channel=1024, bn=True, deform=False):
        super(SSD4Scale, self).__init__()
        self.num_classes, self.size, self.phase, self.bn, self.deform = num_classes, size, phase, bn, deform
        self.backbone = nn.ModuleList(vgg(vgg_base['320'], 3, batch_norm=bn, pool5_ds=True, c7_channel=c7_channel))
        self.L2Norm_4_3 = L2Norm(512, 10)
        self.L2Norm_5_3 = L2Norm(512, 8)
        add_vgg_extras(self, bn, c7_channel)
        self._add_heads([512, 512, c7_channel, 512], num_classes, deform)

    def _sources(self, E, x):
        return E.vgg_trunk(x, self.bn)


def build_net(phase, size=320, num_classes=21, c7_channel=1024, bn=False, deform=False):
    if size not in [320, 512]:
        print("Error: Sorry only SSD320 and SSD512 is supported currently!")
        return
    return SSD4Scale(size, num_classes=num_classes, phase=phase, c7_channel=c7_channel, bn=bn, deform=deform)
