"""Functional CPU-PyTorch restatement of the reference detectors.  TEST INFRASTRUCTURE ONLY.

Restates (state-dict key names are the reference's, SURVEY.md section 5 "Checkpoint"):
  * vgg()                      model/networks.py:136-163
  * conv_dw()                  model/networks.py:736-745
  * L2Norm.forward             layers/modules/l2norm.py:17-21
  * DualRefineDet-VGG(BN)      model/dualrefinedet_vggbn.py:10-117 (params), :119-206 (forward)
  * DualRefineDet-MobileNet    model/dualrefinedet_mobilenet.py:10-125, :127-199
  * RefineDet-VGG              model/refinedet_vgg.py:26-107, :109-219
  * SSD4Scale (TDRN nets)      model/ssd4scale_vgg.py:9-69, :71-135

``param_spec_*`` enumerate (key, shape, kind) so a deterministic random ``state_dict`` can be
generated from a seed on any machine (the 34M-parameter dict is too large to commit);
tests/test_oracle_vs_reference.py loads that dict with strict=True into the *real* reference
modules (build container only) and checks that this restatement reproduces their outputs.
"""
import math

import torch
import torch.nn.functional as F

from .deform_conv_ref import deform_conv_forward

VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512]  # networks.py:15-16
BN_EPS = 1e-5
NUM_BOX = 3


# ----------------------------------------------------------------------------------------------
# parameter specs
# ----------------------------------------------------------------------------------------------
def _conv(spec, name, cout, cin, k, bias=True, kind='conv'):
    spec.append((name + '.weight', (cout, cin, k, k), kind))
    if bias:
        spec.append((name + '.bias', (cout,), 'bias'))


def _bn(spec, name, c):
    spec.append((name + '.weight', (c,), 'bn_weight'))
    spec.append((name + '.bias', (c,), 'bn_bias'))
    spec.append((name + '.running_mean', (c,), 'bn_mean'))
    spec.append((name + '.running_var', (c,), 'bn_var'))
    spec.append((name + '.num_batches_tracked', (), 'bn_count'))


def vgg_layers(bn, c7_channel=1024):
    """Module list layout of vgg() (networks.py:136-163): list of (index, op, args)."""
    layers, idx, cin = [], 0, 3
    for v in VGG_CFG:
        if v == 'M' or v == 'C':
            layers.append((idx, 'pool', {'ceil': v == 'C'}))
            idx += 1
        else:
            layers.append((idx, 'conv', {'cin': cin, 'cout': v, 'k': 3, 'pad': 1, 'dil': 1}))
            idx += 1
            if bn:
                layers.append((idx, 'bn', {'c': v}))
                idx += 1
            layers.append((idx, 'relu', {}))
            idx += 1
            cin = v
    layers.append((idx, 'pool', {'ceil': False}))             # pool5_ds=True: 2x2 s2
    idx += 1
    for (ci, co, k, pad, dil) in ((512, 1024, 3, 6, 6), (1024, c7_channel, 1, 0, 1)):
        layers.append((idx, 'conv', {'cin': ci, 'cout': co, 'k': k, 'pad': pad, 'dil': dil}))
        idx += 1
        if bn:
            layers.append((idx, 'bn', {'c': co}))
            idx += 1
        layers.append((idx, 'relu', {}))
        idx += 1
    return layers


def _spec_vgg_trunk(spec, bn, c7_channel):
    for idx, op, a in vgg_layers(bn, c7_channel):
        if op == 'conv':
            # conv7 gets a damped init so the un-normalised ARM sources stay O(1) (see make_state_dict)
            kind = ('conv_c7_bn' if bn else 'conv_c7') if a['k'] == 1 else 'conv'
            _conv(spec, 'backbone.%d' % idx, a['cout'], a['cin'], a['k'], kind=kind)
        elif op == 'bn':
            _bn(spec, 'backbone.%d' % idx, a['c'])
    spec.append(('L2Norm_4_3.weight', (512,), 'l2norm10'))
    spec.append(('L2Norm_5_3.weight', (512,), 'l2norm8'))


def _spec_vgg_extras(spec, bn, c7_channel):
    if bn:
        _conv(spec, 'extras.0', 256, c7_channel, 1); _bn(spec, 'extras.1', 256)
        _conv(spec, 'extras.3', 512, 256, 3); _bn(spec, 'extras.4', 512)
    else:
        _conv(spec, 'extras.0', 256, c7_channel, 1)
        _conv(spec, 'extras.2', 512, 256, 3)


def _spec_fpn(spec, src_channels, bias):
    _conv(spec, 'last_layer_trans.0', 256, 512, 3, bias)
    _conv(spec, 'last_layer_trans.2', 256, 256, 3, bias)
    _conv(spec, 'last_layer_trans.3', 256, 256, 3, bias)
    for k in range(3):
        _conv(spec, 'trans_layers.%d.0' % k, 256, src_channels[k], 3, bias)
        _conv(spec, 'trans_layers.%d.2' % k, 256, 256, 3, bias)
    for k in range(3):
        spec.append(('up_layers.%d.weight' % k, (256, 256, 2, 2), 'conv'))   # ConvTranspose2d [Cin,Cout,2,2]
        if bias:
            spec.append(('up_layers.%d.bias' % k, (256,), 'bias'))
    for k in range(3):
        _conv(spec, 'latent_layers.%d' % k, 256, 256, 3, bias)


def param_spec_drn_vgg(num_classes=21, c7_channel=1024, def_groups=1, bn=True, multihead=False):
    """dualrefinedet_vggbn.py:22-114."""
    spec = []
    _spec_vgg_trunk(spec, bn, c7_channel)
    src = [512, 512, c7_channel, 512]
    _spec_fpn(spec, src, True)
    _spec_vgg_extras(spec, bn, c7_channel)
    for k in range(4):
        _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3)
    for k in range(4):
        _conv(spec, 'offset.%d' % k, def_groups * 18, NUM_BOX * 4, 1, kind='offset_conv')
    for k in range(4):
        _conv(spec, 'odm_loc.%d' % k, NUM_BOX * 4, 256, 3, False, 'deform')
    for k in range(4):
        _conv(spec, 'odm_conf.%d' % k, NUM_BOX * num_classes, 256, 3, False, 'deform')
    if multihead:
        for k in range(4):
            _conv(spec, 'offset2.%d' % k, def_groups * 50, NUM_BOX * 4, 1, kind='offset_conv')
        for k in range(4):
            _conv(spec, 'odm_loc_2.%d' % k, NUM_BOX * 4, 256, 5, False, 'deform')
        for k in range(4):
            _conv(spec, 'odm_conf_2.%d' % k, NUM_BOX * num_classes, 256, 5, False, 'deform')
    return spec


MOBILENET_DW = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 1), (256, 256, 1), (256, 512, 2),
                (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1),
                (512, 1024, 2), (1024, 1024, 1)]               # dualrefinedet_mobilenet.py:23-35


def _spec_conv_dw(spec, name, inp, oup):
    """conv_dw (networks.py:736-745): indices 0 dw conv, 1 bn, 3 pw conv, 4 bn."""
    spec.append((name + '.0.weight', (inp, 1, 3, 3), 'conv_dw'))
    _bn(spec, name + '.1', inp)
    spec.append((name + '.3.weight', (oup, inp, 1, 1), 'conv'))
    _bn(spec, name + '.4', oup)


def param_spec_drn_mobilenet(num_classes=21, def_groups=1, multihead=False):
    """dualrefinedet_mobilenet.py:19-121 (every conv bias-free)."""
    spec = []
    spec.append(('backbone.0.0.weight', (32, 3, 3, 3), 'conv'))
    _bn(spec, 'backbone.0.1', 32)
    for n, (i, o, s) in enumerate(MOBILENET_DW):
        _spec_conv_dw(spec, 'backbone.%d' % (n + 1), i, o)
    spec.append(('L2Norm_4_3.weight', (512,), 'l2norm20'))
    spec.append(('L2Norm_5_3.weight', (1024,), 'l2norm8'))
    for e, cin in enumerate((1024, 512)):
        _conv(spec, 'extras.%d.0' % e, 256, cin, 1)              # NB: these two 1x1 convs keep bias
        _bn(spec, 'extras.%d.1' % e, 256)
        _spec_conv_dw(spec, 'extras.%d.3' % e, 256, 512)
    src = [512, 1024, 512, 512]
    _spec_fpn(spec, src, False)
    for k in range(4):
        _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3, False)
    for k in range(4):
        _conv(spec, 'offset.%d' % k, def_groups * 18, NUM_BOX * 4, 1, False, 'offset_conv')
    for k in range(4):
        _conv(spec, 'odm_loc.%d' % k, NUM_BOX * 4, 256, 3, False, 'deform')
    for k in range(4):
        _conv(spec, 'odm_conf.%d' % k, NUM_BOX * num_classes, 256, 3, False, 'deform')
    if multihead:
        for k in range(4):
            _conv(spec, 'offset2.%d' % k, def_groups * 50, NUM_BOX * 4, 1, False, 'offset_conv')
        for k in range(4):
            _conv(spec, 'odm_loc_2.%d' % k, NUM_BOX * 4, 256, 5, False, 'deform')
        for k in range(4):
            _conv(spec, 'odm_conf_2.%d' % k, NUM_BOX * num_classes, 256, 5, False, 'deform')
    return spec


def param_spec_refinedet_vgg(num_classes=21, use_refine=True, c7_channel=1024, bn=False, multihead=False):
    """refinedet_vgg.py:36-104."""
    spec = []
    _spec_vgg_trunk(spec, bn, c7_channel)
    src = [512, 512, c7_channel, 512]
    _spec_fpn(spec, src, True)
    _spec_vgg_extras(spec, bn, c7_channel)
    if use_refine:
        for k in range(4):
            _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3)
    for k in range(4):
        _conv(spec, 'odm_loc.%d' % k, NUM_BOX * 4, 256, 3)
    for k in range(4):
        _conv(spec, 'odm_conf.%d' % k, NUM_BOX * num_classes, 256, 3)
    if multihead:
        for k in range(4):
            _conv(spec, 'odm_loc_2.%d' % k, NUM_BOX * 4, 256, 5)
        for k in range(4):
            _conv(spec, 'odm_conf_2.%d' % k, NUM_BOX * num_classes, 256, 5)
    return spec


def param_spec_ssd4scale_vgg(num_classes=31, c7_channel=1024, bn=True, deform=False):
    """ssd4scale_vgg.py:18-66."""
    spec = []
    _spec_vgg_trunk(spec, bn, c7_channel)
    _spec_vgg_extras(spec, bn, c7_channel)
    src = [512, 512, c7_channel, 512]
    if deform:
        for k in range(4):
            _conv(spec, 'offset.%d' % k, 8 * 18, NUM_BOX * 4, 1, kind='offset_conv')
        for k in range(4):
            _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3, False, 'deform')
        for k in range(4):
            _conv(spec, 'arm_conf.%d' % k, NUM_BOX * num_classes, src[k], 3, False, 'deform')
    else:
        for k in range(4):
            _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3)
        for k in range(4):
            _conv(spec, 'arm_conf.%d' % k, NUM_BOX * num_classes, src[k], 3)
    return spec


def param_spec_ssd4scale_mobile(num_classes=31, c7_channel=1024, deform=False):
    """ssd4scale_mobile.py:19-82 (backbone as MOBILENET_DW with the last block 1024 -> c7_channel; heads keep bias
    when plain, are bias-free ConvOffset2d with 8 deformable groups when ``deform``)."""
    spec = []
    spec.append(('backbone.0.0.weight', (32, 3, 3, 3), 'conv'))
    _bn(spec, 'backbone.0.1', 32)
    for n, (i, o, s) in enumerate(MOBILENET_DW):
        _spec_conv_dw(spec, 'backbone.%d' % (n + 1), i, c7_channel if n == len(MOBILENET_DW) - 1 else o)
    spec.append(('L2Norm_4_3.weight', (512,), 'l2norm10'))       # :40 (scale 10; the DualRefineDet variant uses 20)
    spec.append(('L2Norm_5_3.weight', (1024,), 'l2norm8'))       # :41
    for e, cin in enumerate((c7_channel, 512)):                  # :43-51
        _conv(spec, 'extras.%d.0' % e, 256, cin, 1)
        _bn(spec, 'extras.%d.1' % e, 256)
        _spec_conv_dw(spec, 'extras.%d.3' % e, 256, 512)
    src = [512, c7_channel, 512, 512]
    if deform:                                                   # :52-71
        for k in range(4):
            _conv(spec, 'offset.%d' % k, 8 * 18, NUM_BOX * 4, 1, kind='offset_conv')
        for k in range(4):
            _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3, False, 'deform')
        for k in range(4):
            _conv(spec, 'arm_conf.%d' % k, NUM_BOX * num_classes, src[k], 3, False, 'deform')
    else:                                                        # :72-82
        for k in range(4):
            _conv(spec, 'arm_loc.%d' % k, NUM_BOX * 4, src[k], 3)
        for k in range(4):
            _conv(spec, 'arm_conf.%d' % k, NUM_BOX * num_classes, src[k], 3)
    return spec


def make_state_dict(spec, seed=0, offset_gain=1.5):
    """Deterministic random weights (CPU mt19937): He-uniform convs, randomised BN statistics so
    BN folding is exercised, xavier-uniform deformable weights (networks.py:727)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    for name, shape, kind in spec:
        if kind in ('conv', 'conv_dw', 'offset_conv', 'conv_c7', 'conv_c7_bn'):
            fan_in = shape[1] * shape[2] * shape[3]
            gain = {'offset_conv': offset_gain, 'conv_c7': 0.2, 'conv_c7_bn': 0.04}.get(kind, 1.0)
            sd[name] = uni(shape, gain * math.sqrt(6.0 / fan_in))
        elif kind == 'deform':
            fan_in = shape[1] * shape[2] * shape[3]
            fan_out = shape[0] * shape[2] * shape[3]
            sd[name] = uni(shape, math.sqrt(6.0 / (fan_in + fan_out)))
        elif kind == 'bias':
            sd[name] = uni(shape, 0.1)
        elif kind == 'bn_weight':
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif kind == 'bn_bias' or kind == 'bn_mean':
            sd[name] = torch.randn(shape, generator=g) * 0.1
        elif kind == 'bn_var':
            sd[name] = torch.rand(shape, generator=g) + 0.5
        elif kind == 'bn_count':
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif kind.startswith('l2norm'):
            sd[name] = torch.full(shape, float(kind[6:])) * (0.9 + 0.2 * torch.rand(shape, generator=g))
        else:
            raise KeyError(kind)
    return sd


def state_dict_checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if v.is_floating_point()))


# ----------------------------------------------------------------------------------------------
# functional forwards
# ----------------------------------------------------------------------------------------------
def _c(sd, name, x, stride=1, pad=0, dil=1, groups=1):
    return F.conv2d(x, sd[name + '.weight'], sd.get(name + '.bias'), stride, pad, dil, groups)


def _b(sd, name, x):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], False, 0.0, BN_EPS)


def l2norm(x, weight, eps=1e-10):
    """l2norm.py:17-21."""
    norm = x.pow(2).sum(dim=1, keepdim=True).sqrt() + eps
    return weight.view(1, -1, 1, 1) * (x / norm)


def _vgg_trunk(sd, x, bn, c7_channel=1024):
    """dualrefinedet_vggbn.py:130-153: returns the four ARM sources."""
    layers = vgg_layers(bn, c7_channel)
    split43, split53 = (23, 33)[bn], (30, 43)[bn]
    sources = []
    for idx, op, a in layers:
        if idx == split43:
            sources.append(l2norm(x, sd['L2Norm_4_3.weight']))
        if idx == split53:
            sources.append(l2norm(x, sd['L2Norm_5_3.weight']))
        if op == 'conv':
            x = _c(sd, 'backbone.%d' % idx, x, 1, a['pad'], a['dil'])
        elif op == 'bn':
            x = _b(sd, 'backbone.%d' % idx, x)
        elif op == 'relu':
            x = F.relu(x)
        elif op == 'pool':
            x = F.max_pool2d(x, 2, 2, ceil_mode=a['ceil'])
    sources.append(x)
    if bn:
        x = F.relu(_b(sd, 'extras.1', _c(sd, 'extras.0', x)))
        x = F.relu(_b(sd, 'extras.4', _c(sd, 'extras.3', x, 2, 1)))
    else:
        x = F.relu(_c(sd, 'extras.0', x))
        x = F.relu(_c(sd, 'extras.2', x, 2, 1))
    sources.append(x)
    return sources


def _fpn(sd, arm_sources):
    """TCB / top-down path, dualrefinedet_vggbn.py:166-179."""
    x = arm_sources[3]
    x = _c(sd, 'last_layer_trans.0', x, 1, 1)
    x = F.relu(x)
    x = _c(sd, 'last_layer_trans.2', x, 1, 1)
    x = _c(sd, 'last_layer_trans.3', x, 1, 1)
    odm_sources = [x]
    trans = []
    for k in range(3):
        t = F.relu(_c(sd, 'trans_layers.%d.0' % k, arm_sources[k], 1, 1))
        trans.append(_c(sd, 'trans_layers.%d.2' % k, t, 1, 1))
    trans.reverse()
    for k in range(3):
        up = F.conv_transpose2d(x, sd['up_layers.%d.weight' % k], sd.get('up_layers.%d.bias' % k), 2, 0)
        x = F.relu(_c(sd, 'latent_layers.%d' % k, F.relu(up + trans[k]), 1, 1))
        odm_sources.append(x)
    odm_sources.reverse()
    return odm_sources


def _flat(t):
    return t.permute(0, 2, 3, 1).contiguous().view(t.size(0), -1)


def _drn_heads(sd, arm_sources, odm_sources, num_classes, dg, multihead, softmax):
    """dualrefinedet_vggbn.py:154-197 (identical in dualrefinedet_mobilenet.py:165-188)."""
    arm_loc, offsets, offsets2 = [], [], []
    for k in range(4):
        loc_a = _c(sd, 'arm_loc.%d' % k, arm_sources[k], 1, 1)
        arm_loc.append(_flat(loc_a))
        offsets.append(_c(sd, 'offset.%d' % k, loc_a))
        if multihead:
            offsets2.append(_c(sd, 'offset2.%d' % k, loc_a))
    odm_loc, odm_conf = [], []
    for k in range(4):
        ob = odm_sources[k]
        l = deform_conv_forward(ob, offsets[k], sd['odm_loc.%d.weight' % k], 1, 1, 1, dg)
        c = deform_conv_forward(ob, offsets[k], sd['odm_conf.%d.weight' % k], 1, 1, 1, dg)
        if multihead:
            l = l + deform_conv_forward(ob, offsets2[k], sd['odm_loc_2.%d.weight' % k], 1, 2, 1, dg)
            c = c + deform_conv_forward(ob, offsets2[k], sd['odm_conf_2.%d.weight' % k], 1, 2, 1, dg)
        odm_loc.append(_flat(l))
        odm_conf.append(_flat(c))
    arm_loc = torch.cat(arm_loc, 1)
    odm_loc = torch.cat(odm_loc, 1)
    odm_conf = torch.cat(odm_conf, 1)
    b = arm_loc.size(0)
    conf = odm_conf.view(-1, num_classes)
    if softmax:
        conf = F.softmax(conf, dim=1)
    return arm_loc.view(b, -1, 4), offsets, odm_loc.view(b, -1, 4), conf


def drn_vgg_forward(sd, x, num_classes=21, c7_channel=1024, def_groups=1, bn=True, multihead=False,
                    softmax=True):
    with torch.no_grad():
        arm_sources = _vgg_trunk(sd, x, bn, c7_channel)
        odm_sources = _fpn(sd, arm_sources)
        return _drn_heads(sd, arm_sources, odm_sources, num_classes, def_groups, multihead, softmax)


def _conv_dw(sd, name, x, stride):
    c = x.size(1)
    x = F.relu(_b(sd, name + '.1', _c(sd, name + '.0', x, stride, 1, 1, c)))
    return F.relu(_b(sd, name + '.4', _c(sd, name + '.3', x)))


def _mobilenet_trunk(sd, x):
    """The four ARM sources of the MobileNet trunk (dualrefinedet_mobilenet.py:139-152)."""
    x = F.relu(_b(sd, 'backbone.0.1', _c(sd, 'backbone.0.0', x, 2, 1)))
    arm_sources = []
    for n, (i, o, s) in enumerate(MOBILENET_DW):
        if n + 1 == 12:
            arm_sources.append(l2norm(x, sd['L2Norm_4_3.weight']))
        x = _conv_dw(sd, 'backbone.%d' % (n + 1), x, s)
    arm_sources.append(l2norm(x, sd['L2Norm_5_3.weight']))
    for e in range(2):
        x = F.relu(_b(sd, 'extras.%d.1' % e, _c(sd, 'extras.%d.0' % e, x)))
        x = _conv_dw(sd, 'extras.%d.3' % e, x, 2)
        arm_sources.append(x)
    return arm_sources


def drn_mobilenet_forward(sd, x, num_classes=21, def_groups=1, multihead=False, softmax=True):
    """dualrefinedet_mobilenet.py:127-199; slot 1 of the reference output is None (:188)."""
    with torch.no_grad():
        arm_sources = _mobilenet_trunk(sd, x)
        odm_sources = _fpn(sd, arm_sources)
        return _drn_heads(sd, arm_sources, odm_sources, num_classes, def_groups, multihead, softmax)


def refinedet_vgg_forward(sd, x, num_classes=21, use_refine=True, c7_channel=1024, bn=False,
                          multihead=False, softmax=True):
    """refinedet_vgg.py:109-219 (plain-conv ODM heads, BASELINE config 1)."""
    with torch.no_grad():
        arm_sources = _vgg_trunk(sd, x, bn, c7_channel)
        arm_loc = None
        if use_refine:
            arm_loc = torch.cat([_flat(_c(sd, 'arm_loc.%d' % k, arm_sources[k], 1, 1)) for k in range(4)], 1)
        odm_sources = _fpn(sd, arm_sources)
        ls, cs = [], []
        for k in range(4):
            l = _c(sd, 'odm_loc.%d' % k, odm_sources[k], 1, 1)
            c = _c(sd, 'odm_conf.%d' % k, odm_sources[k], 1, 1)
            if multihead:
                l = l + _c(sd, 'odm_loc_2.%d' % k, odm_sources[k], 1, 2)
                c = c + _c(sd, 'odm_conf_2.%d' % k, odm_sources[k], 1, 2)
            ls.append(_flat(l)); cs.append(_flat(c))
        odm_loc = torch.cat(ls, 1)
        conf = torch.cat(cs, 1).view(-1, num_classes)
        if softmax:
            conf = F.softmax(conf, dim=1)
        b = x.size(0)
        if use_refine:
            return arm_loc.view(b, -1, 4), None, odm_loc.view(b, -1, 4), conf
        return odm_loc.view(b, -1, 4), conf


def ssd4scale_vgg_forward(sd, x, num_classes=31, c7_channel=1024, bn=True, deform=False,
                          ref_loc=(), offset_list=(), ret_loc=False, ret_off=False, softmax=True):
    """ssd4scale_vgg.py:71-135."""
    with torch.no_grad():
        offsets = None
        if deform:
            if not offset_list:
                offsets = [_c(sd, 'offset.%d' % k, ref_loc[k]) for k in range(4)]   # :72-76
            else:
                offsets = list(offset_list)
        src = _vgg_trunk(sd, x, bn, c7_channel)
        return _ssd4scale_heads(sd, src, x, num_classes, deform, offsets, ret_loc, ret_off, softmax)


def _ssd4scale_heads(sd, src, x, num_classes, deform, offsets, ret_loc, ret_off, softmax):
    """Head loop and output tuple shared by the two SSD4Scale variants (ssd4scale_vgg.py:104-135,
    ssd4scale_mobile.py:108-140)."""
    locs, confs, loc_maps = [], [], []
    for k in range(4):
        if deform:
            l = deform_conv_forward(src[k], offsets[k], sd['arm_loc.%d.weight' % k], 1, 1, 1, 8)
            c = deform_conv_forward(src[k], offsets[k], sd['arm_conf.%d.weight' % k], 1, 1, 1, 8)
        else:
            l = _c(sd, 'arm_loc.%d' % k, src[k], 1, 1)
            c = _c(sd, 'arm_conf.%d' % k, src[k], 1, 1)
            loc_maps.append(l)
        locs.append(_flat(l)); confs.append(_flat(c))
    loc = torch.cat(locs, 1)
    conf = torch.cat(confs, 1).view(-1, num_classes)
    if softmax:
        conf = F.softmax(conf, dim=1)
    out = [loc.view(x.size(0), -1, 4), conf]
    if ret_loc:
        out.append(loc_maps)
    if ret_off:
        out.append(offsets)
    return tuple(out)


def ssd4scale_mobile_forward(sd, x, num_classes=31, c7_channel=1024, deform=False,
                             ref_loc=(), offset_list=(), ret_loc=False, ret_off=False, softmax=True):
    """ssd4scale_mobile.py:86-140."""
    with torch.no_grad():
        offsets = None
        if deform:
            if not offset_list:
                offsets = [_c(sd, 'offset.%d' % k, ref_loc[k]) for k in range(4)]   # :87-93
            else:
                offsets = list(offset_list)
        inp = x
        x = F.relu(_b(sd, 'backbone.0.1', _c(sd, 'backbone.0.0', x, 2, 1)))
        src = []
        for n, (i, o, s) in enumerate(MOBILENET_DW):
            if n + 1 == 12:                                                         # :101-103
                src.append(l2norm(x, sd['L2Norm_4_3.weight']))
            x = _conv_dw(sd, 'backbone.%d' % (n + 1), x, s)
        src.append(l2norm(x, sd['L2Norm_5_3.weight']))                              # :104-106
        for e in range(2):                                                          # :108-110
            x = F.relu(_b(sd, 'extras.%d.1' % e, _c(sd, 'extras.%d.0' % e, x)))
            x = _conv_dw(sd, 'extras.%d.3' % e, x, 2)
            src.append(x)
        return _ssd4scale_heads(sd, src, inp, num_classes, deform, offsets, ret_loc, ret_off, softmax)
