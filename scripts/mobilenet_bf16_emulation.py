"""CPU emulation: which roundings of the bf16 MobileNet trunk cost what on the ODM outputs (reference offsets injected)?

The DualRefineDet-MobileNet trunk is 1 + 13 x (depthwise 3x3, pointwise 1x1) + extras; the bf16 path rounds every stored
activation and every tensor-core operand to bf16.  This script re-evaluates the oracle's forward (oracle/model_ref.py) with
chosen subsets of those roundings and reports max |err| / max |ref| on odm_loc / conf (the metric of
tests/test_gpu_models.py::test_bf16_heads_with_reference_offsets), so that a fused depthwise->pointwise kernel can be judged
before it is written:
  all        every dw output, pw output and pw weight rounded (what the two-kernel path does)
  fused      as `all` (a fused kernel still rounds the dw output once to make the MMA operand)
  a16        dw output kept at 16 mantissa bits (hi + lo operand, two MMAs per k-block)
  a16w16     + pointwise weights at 16 mantissa bits (three products)
  store      only the stored pointwise outputs rounded (the floor of any scheme that keeps bf16 activations in HBM)
Test infrastructure only (imports oracle/): never imported by the product.
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import model_ref as M                      # noqa: E402
from oracle.make_golden import CASES, SEED_W, make_input       # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).float()


def r16(t):
    hi = bf(t)
    return hi + bf(t - hi)


def fold(sd, conv, bn):
    w = sd[conv + '.weight'].double()
    g, b, m, v = [sd[bn + k].double() for k in ('.weight', '.bias', '.running_mean', '.running_var')]
    s = g / torch.sqrt(v + M.BN_EPS)
    cb = sd[conv + '.bias'].double() if conv + '.bias' in sd else 0.0
    return (w * s.view(-1, 1, 1, 1)).float(), (b + (cb - m) * s).float()


def trunk(sd, x, r_dw, r_pw, r_w):
    """r_dw / r_pw / r_w: rounding applied to dw outputs, pw outputs (stored activations), pw weights."""
    w, b = fold(sd, 'backbone.0.0', 'backbone.0.1')
    x = r_pw(F.relu(F.conv2d(x, w, b, 2, 1)))          # stem conv (CUDA cores, fp32 weights), stored bf16
    src = []

    def conv_dw(name, x, stride):
        c = x.size(1)
        w, b = fold(sd, name + '.0', name + '.1')
        x = r_dw(F.relu(F.conv2d(x, w, b, stride, 1, 1, c)))
        w, b = fold(sd, name + '.3', name + '.4')
        return r_pw(F.relu(F.conv2d(x, r_w(w), b)))

    for n, (i, o, s) in enumerate(M.MOBILENET_DW):
        if n + 1 == 12:
            src.append(r_pw(M.l2norm(x, sd['L2Norm_4_3.weight'])))
        x = conv_dw('backbone.%d' % (n + 1), x, s)
    src.append(r_pw(M.l2norm(x, sd['L2Norm_5_3.weight'])))
    for e in range(2):
        w, b = fold(sd, 'extras.%d.0' % e, 'extras.%d.1' % e)
        x = r_pw(F.relu(F.conv2d(x, r_w(w), b)))
        x = conv_dw('extras.%d.3' % e, x, 2)
        src.append(x)
    return src


def main():
    mod_name, spec_fn, build_kw, spec_kw, _ = CASES['drn_mobilenet320']
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    x = make_input(2, 320, seed=9)
    ident = lambda t: t
    with torch.no_grad():
        ref = M.drn_mobilenet_forward(sd, x, **spec_kw)
        offs = ref[1]

        def heads(src):
            odm = M._fpn(sd, src)
            loc, conf = [], []
            for k in range(4):
                loc.append(M._flat(M.deform_conv_forward(odm[k], offs[k], sd['odm_loc.%d.weight' % k], 1, 1, 1, spec_kw.get('def_groups', 1))))
                conf.append(M._flat(M.deform_conv_forward(odm[k], offs[k], sd['odm_conf.%d.weight' % k], 1, 1, 1, spec_kw.get('def_groups', 1))))
            loc = torch.cat(loc, 1).view(x.size(0), -1, 4)
            conf = F.softmax(torch.cat(conf, 1).view(-1, spec_kw.get('num_classes', 21)), 1)
            return loc, conf

        base = heads(trunk(sd, x, ident, ident, ident))
        err = lambda a, b: float((a - b).abs().max() / b.abs().max())
        print('fp32 trunk through this script vs oracle: loc %.2e conf %.2e' % (err(base[0], ref[2]), err(base[1], ref[3])))
        for name, (r_dw, r_pw, r_w) in [('all', (bf, bf, bf)), ('a16', (r16, bf, bf)), ('a16w16', (r16, bf, r16)),
                                        ('store', (ident, bf, ident)), ('w only', (ident, ident, bf)), ('dw only', (bf, ident, ident))]:
            out = heads(trunk(sd, x, r_dw, r_pw, r_w))
            print('%-8s trunk roundings -> odm_loc %.2e  conf %.2e   (FPN and heads in fp32)' % (name, err(out[0], ref[2]), err(out[1], ref[3])))


if __name__ == '__main__':
    main()
