"""DualRefineDet-VGG(BN): drop-in for the reference's model/dualrefinedet_vggbn.py.

Same ``build_net`` signature (:217-222), same state-dict keys (:22-114), same forward outputs
(:191-197): (arm_loc [B,P,4], [offset maps NCHW], odm_loc [B,P,4], softmax(conf) [B*P,C]).
The compute is the C-ABI executor in _engine.py, not nn.Module.__call__ of the children.
"""
import torch
import torch.nn as nn

from .. import ops
from ..layers.modules.l2norm import L2Norm
from ._base import DetectorBase
from ._engine import prior_layout
from .networks import vgg, vgg_base, ConvOffset2d


def _list(fn, n=4):
    return nn.ModuleList([fn(k) for k in range(n)])


def add_fpn(m, src, bias=True):
    """TCB/FPN parameter containers shared by the RefineDet family (dualrefinedet_vggbn.py:30-34,97-114)."""
    c3 = lambda i, o: nn.Conv2d(i, o, kernel_size=3, stride=1, padding=1, bias=bias)
    m.last_layer_trans = nn.Sequential(c3(512, 256), nn.ReLU(inplace=True), c3(256, 256), c3(256, 256))
    m.trans_layers = nn.ModuleList([nn.Sequential(c3(src[k], 256), nn.ReLU(inplace=True), c3(256, 256))
                                    for k in range(3)])
    m.up_layers = _list(lambda k: nn.ConvTranspose2d(256, 256, kernel_size=2, stride=2, padding=0, bias=bias), 3)
    m.latent_layers = _list(lambda k: c3(256, 256), 3)


def add_vgg_extras(m, bn, c7_channel):
    layers = [nn.Conv2d(c7_channel, 256, kernel_size=1)]
    if bn:
        layers.append(nn.BatchNorm2d(256))
    layers += [nn.ReLU(inplace=True), nn.Conv2d(256, 512, kernel_size=3, stride=2, padding=1)]
    if bn:
        layers.append(nn.BatchNorm2d(512))
    layers.append(nn.ReLU(inplace=True))
    m.extras = nn.Sequential(*layers)


def add_deform_heads(m, num_classes, dg, multihead, bias):
    nb = 3
    m.offset = _list(lambda k: nn.Conv2d(nb * 4, dg * 2 * 9, kernel_size=1, bias=bias))
    m.odm_loc = _list(lambda k: ConvOffset2d(256, nb * 4, 3, 1, 1, num_deformable_groups=dg))
    m.odm_conf = _list(lambda k: ConvOffset2d(256, nb * num_classes, 3, 1, 1, num_deformable_groups=dg))
    if multihead:
        m.offset2 = _list(lambda k: nn.Conv2d(nb * 4, dg * 2 * 25, kernel_size=1, bias=bias))
        m.odm_loc_2 = _list(lambda k: ConvOffset2d(256, nb * 4, 5, 1, 2, 1, dg))
        m.odm_conf_2 = _list(lambda k: ConvOffset2d(256, nb * num_classes, 5, 1, 2, 1, dg))


class RefineSSD(DetectorBase):
    def __init__(self, size, num_classes=21, phase='train', c7_channel=1024, def_groups=1, bn=True,
                 multihead=False, return_feature=False, device='cuda'):
        super(RefineSSD, self).__init__()
        self.num_classes, self.size, self.phase = num_classes, size, phase
        self.def_groups, self.bn, self.multihead = def_groups, bn, multihead
        self.return_feature, self.device = return_feature, device
        self.c7_channel = c7_channel
        if return_feature:
            raise NotImplementedError('return_feature (DetectOTA tubelet feature) is outside the hot path')
        self.backbone = nn.ModuleList(vgg(vgg_base['320'], 3, batch_norm=bn, pool5_ds=True, c7_channel=c7_channel))
        self.L2Norm_4_3 = L2Norm(512, 10)
        self.L2Norm_5_3 = L2Norm(512, 8)
        src = [512, 512, c7_channel, 512]
        add_fpn(self, src)
        add_vgg_extras(self, bn, c7_channel)
        self.arm_loc = _list(lambda k: nn.Conv2d(src[k], 12, kernel_size=3, stride=1, padding=1))
        add_deform_heads(self, num_classes, def_groups, multihead, True)
        if phase == 'test':
            self.softmax = nn.Softmax(dim=1)

    def forward(self, x, _offsets=None):
        """``_offsets`` (test hook, not part of the reference signature): (offset_list, offset2_list) of NCHW
        fp32 maps that replace the ARM-regressed offsets, to check the deformable heads in isolation."""
        E = self.engine()
        x = self._check_input(x)
        if x.shape[2] != x.shape[3]:
            raise ValueError('square inputs only (got %dx%d)' % (x.shape[2], x.shape[3]))
        # fully convolutional like the reference's forward: the pyramid follows the INPUT size (multi-scale testing runs one
        # module at several sizes, multi_eval.py:531-556), self.size is only the nominal training size
        arm_loc, offs, offs2, offs_nchw, odm_sources, P, lv = E.trunk_arm_tcb(x, self.bn, int(x.shape[2]), self.multihead)
        if _offsets is not None:
            offs = [ops.nchw_f32_to_nhwc(o.float(), torch.float32) for o in _offsets[0]]
            offs2 = [ops.nchw_f32_to_nhwc(o.float(), torch.float32) for o in (_offsets[1] or [])]
            offs_nchw = [ops.nhwc_to_nchw_f32(o) for o in offs]
        odm_loc, conf = E.deform_heads(odm_sources, offs, offs2, P, lv, self.num_classes, self.def_groups,
                                       self.multihead)
        return arm_loc, offs_nchw, odm_loc, conf


def build_net(phase, size=320, num_classes=21, c7_channel=1024, def_groups=1, bn=True, multihead=False,
              return_feature=False):
    if size not in [320, 512]:
        print("Error: Sorry only SSD320 and SSD512 is supported currently!")
        return
    return RefineSSD(size, num_classes=num_classes, phase=phase, c7_channel=c7_channel, def_groups=def_groups,
                     bn=bn, multihead=multihead, return_feature=return_feature)
