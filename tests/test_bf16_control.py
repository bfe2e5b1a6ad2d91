"""CPU control for the bf16 end-to-end gate on the deformable ODM heads.

Claim under test: "the rows of odm_loc / conf that the bf16 path gets wrong by more than 2e-2 are rows where the
rounding of the ARM regression moved a sampling tap across the edge of the map -- the reference's sampler is
discontinuous there (deform_conv_cuda_kernel.cu:195) -- and the ORACLE ITSELF does the same when its own fp32 ARM
regression is perturbed by that much".  Here the oracle is run against itself, no GPU involved:

  * `bf16-emulated`: the VGG-BN trunk and ARM heads re-evaluated with the roundings of the tensor-core path (BN folded
    in float64, weights and every stored activation rounded to bf16, fp32 accumulation) -> an ARM regression that is
    off by 5-8e-3 max-norm, which is what the B200 shows (scripts/bf16_attribution.py, profiles/r02a_bf16_attribution.txt:
    7.7e-3 / 5.4e-3 / 6.4e-3 / 7.2e-3 per level);
  * `noise`: the fp32 regression plus uniform noise of the same max-norm size (7e-3).

In both cases the fp32 oracle heads, fed the fp32 ODM sources and the offsets regressed from the perturbed ARM maps, are
compared with the clean oracle.  The assertions are the same as on the GPU (tests/test_gpu_parity_attribution.py):
every row further than 2e-2 from the clean oracle has a tap that changed side (for the emulated rounding: all of them;
for white noise, whose RMS is ~3x that of the rounding at the same max-norm: >= 90 %), and every row without such a tap
stays inside a bound that scales with the perturbation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import parity_tools as PT
from oracle import model_ref as M
from oracle.make_golden import CASES, SEED_W, make_input

TOL = 2e-2


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _fold(sd, name, bnname):
    """eval-mode BatchNorm folded into the conv in float64 (what tdrn_b200.ops.PackedConv does)."""
    w = sd[name + '.weight'].double(); b = sd[name + '.bias'].double()
    s = sd[bnname + '.weight'].double() / torch.sqrt(sd[bnname + '.running_var'].double() + M.BN_EPS)
    return (w * s.view(-1, 1, 1, 1)).float(), ((b - sd[bnname + '.running_mean'].double()) * s + sd[bnname + '.bias'].double()).float()


def vgg_trunk_bf16_emulated(sd, x):
    """oracle.model_ref._vgg_trunk (bn=True) with the storage roundings of the bf16 tensor-core path."""
    sources = []
    for idx, op, a in M.vgg_layers(True, 1024):
        if idx == 33:
            sources.append(_bf(M.l2norm(x, sd['L2Norm_4_3.weight'])))
        if idx == 43:
            sources.append(_bf(M.l2norm(x, sd['L2Norm_5_3.weight'])))
        if op == 'conv':
            w, b = _fold(sd, 'backbone.%d' % idx, 'backbone.%d' % (idx + 1))
            xin = x if idx == 0 else _bf(x)                   # the stem reads the fp32 image
            x = _bf(F.relu(F.conv2d(xin, w if idx == 0 else _bf(w), b, 1, a['pad'], a['dil'])))
        elif op == 'pool':
            x = F.max_pool2d(x, 2, 2, ceil_mode=a['ceil'])
    sources.append(x)
    w, b = _fold(sd, 'extras.0', 'extras.1'); x = _bf(F.relu(F.conv2d(x, _bf(w), b)))
    w, b = _fold(sd, 'extras.3', 'extras.4'); x = _bf(F.relu(F.conv2d(x, _bf(w), b, 2, 1)))
    sources.append(x)
    return sources


@pytest.fixture(scope='module')
def clean():
    torch.manual_seed(0)
    mod_name, spec_fn, build_kw, spec_kw, _ = CASES['drn_vgg320_multihead']
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    x = make_input(1, 320)
    with torch.no_grad():
        src = M._vgg_trunk(sd, x, True)
        odm = M._fpn(sd, src)
        loc_a = [M._c(sd, 'arm_loc.%d' % k, src[k], 1, 1) for k in range(4)]

        def heads(loc_maps):
            o1 = [M._c(sd, 'offset.%d' % k, loc_maps[k]) for k in range(4)]
            o2 = [M._c(sd, 'offset2.%d' % k, loc_maps[k]) for k in range(4)]
            l, c = PT.odm_heads_from_offsets(sd, odm, o1, o2, 21)
            return o1, o2, l[0].numpy(), c.numpy()

        o1, o2, l, c = heads(loc_a)
        # the helper is the oracle's own head loop: same numbers as the whole-model restatement
        ref = M.drn_vgg_forward(sd, x, **spec_kw)
        assert np.array_equal(l, ref[2][0].numpy()) and np.array_equal(c, ref[3].numpy())
    return dict(sd=sd, x=x, loc_a=loc_a, heads=heads, o1=o1, o2=o2, l=l, c=c,
                amax=max(float(m.abs().max()) for m in loc_a))


def _report(clean, loc_maps):
    with torch.no_grad():
        p1, p2, l, c = clean['heads'](loc_maps)
    fl = [PT.flipped_pixels(clean['o1'][k], p1[k], 3, 1, 1) | PT.flipped_pixels(clean['o2'][k], p2[k], 5, 2, 1) for k in range(4)]
    rows = PT.flipped_rows(fl)[0]
    return PT.split_report(l, clean['l'], rows, TOL), PT.split_report(c, clean['c'], rows, TOL)


def test_oracle_with_bf16_emulated_arm_regression(clean):
    sd, x = clean['sd'], clean['x']
    with torch.no_grad():
        srcb = vgg_trunk_bf16_emulated(sd, x)
        loc_b = [F.conv2d(srcb[k], _bf(sd['arm_loc.%d.weight' % k]), sd['arm_loc.%d.bias' % k], 1, 1) for k in range(4)]
    errs = [float((loc_b[k] - clean['loc_a'][k]).abs().max()) / clean['amax'] for k in range(4)]
    # the perturbation is the size the B200 shows (5.4-7.7e-3) and inside the 2e-2 the ARM tensors are held to
    assert 3e-3 < max(errs) < 1.2e-2, errs
    for rep in _report(clean, loc_b):
        # the oracle itself, fed an ARM regression that is within 2e-2, lands outside 2e-2 on some rows ...
        assert rep['n_beyond'] >= 10, rep
        # ... every one of them has a tap that changed side of the map edge ...
        assert rep['n_beyond_flipped'] == rep['n_beyond'], rep
        # ... and all other rows stay inside the bound; whole-tensor L2 ~1-1.6e-2, dominated by the flipped rows
        assert rep['max_other'] < TOL, rep
        assert rep['l2'] < 3e-2, rep
        assert rep['n_flipped'] < 0.03 * rep['rows'], rep


@pytest.mark.parametrize('seed', [100, 101])
def test_oracle_with_noise_of_the_same_size(clean, seed):
    g = torch.Generator().manual_seed(seed)
    eps = 7e-3
    pert = [m + (torch.rand(m.shape, generator=g) * 2 - 1) * eps * clean['amax'] for m in clean['loc_a']]
    for rep in _report(clean, pert):
        assert rep['n_beyond'] >= 50, rep
        assert rep['n_beyond_flipped'] >= 0.9 * rep['n_beyond'], rep       # white noise: RMS ~3x the rounding's
        assert rep['max_other'] < 2 * TOL, rep
        assert rep['l2'] < 5e-2, rep


def test_tap_validity_matches_the_oracle_sampler():
    """tests/parity_tools.tap_validity restates deform_conv_cuda_kernel.cu:195; check it against the oracle im2col:
    a tap reported invalid contributes exactly zero columns for an all-ones input, a valid one a non-zero value."""
    from oracle.deform_conv_ref import deform_im2col
    g = torch.Generator().manual_seed(3)
    for k, pad, dg, c in ((3, 1, 1, 2), (5, 2, 1, 1), (3, 1, 2, 4)):
        h, w = 7, 9
        off = torch.randn(1, dg * 2 * k * k, h, w, generator=g) * 2.5
        cols = deform_im2col(torch.ones(c, h, w), off[0], k, k, 1, pad, 1, dg).view(c, k * k, h, w)
        valid = PT.tap_validity(off, k, pad, dg)[0].view(dg, k * k, h, w)
        cpg = c // dg
        for ch in range(c):
            assert torch.equal(cols[ch] != 0, valid[ch // cpg])


# ---------------------------------------------------------------------------------------------------------------
# MobileNet variant (16-bit path = IEEE-half trunk + bf16 ARM heads / TCB / deformable heads): the same control.
# The emulation below is the product's arithmetic plan (tdrn_b200/model/dualrefinedet_mobilenet.py:mobilenet_sources):
# half image / weights / activations through the first extras block, packed-half FMAs in the depthwise convs (every FMA
# rounded to half), bf16 where the sources leave the trunk, bf16 second extras block, bf16 ARM heads.
# ---------------------------------------------------------------------------------------------------------------
def _h(t):
    return t.to(torch.float16).to(torch.float32)


def _fold_nb(sd, conv, bn):
    w = sd[conv + '.weight'].double()
    g, b, m, v = [sd[bn + k].double() for k in ('.weight', '.bias', '.running_mean', '.running_var')]
    s = g / torch.sqrt(v + M.BN_EPS)
    cb = sd[conv + '.bias'].double() if conv + '.bias' in sd else 0.0
    return (w * s.view(-1, 1, 1, 1)).float(), (b + (cb - m) * s).float()


def _dw_packed_half(x, w, b, stride):
    c = x.size(1)
    xp = F.pad(x, (1, 1, 1, 1))
    Ho, Wo = (x.shape[2] - 1) // stride + 1, (x.shape[3] - 1) // stride + 1
    acc = _h(b).view(1, c, 1, 1).expand(x.size(0), c, Ho, Wo).clone()
    wh = _h(w)
    for i in range(3):
        for j in range(3):
            xt = xp[:, :, i:i + (Ho - 1) * stride + 1:stride, j:j + (Wo - 1) * stride + 1:stride]
            acc = _h(acc + xt * wh[:, 0, i, j].view(1, c, 1, 1))          # fma.rn.f16x2: one rounding per tap
    return F.relu(acc)


def mobilenet_trunk_half_emulated(sd, x):
    w, b = _fold_nb(sd, 'backbone.0.0', 'backbone.0.1')
    x = _h(F.relu(F.conv2d(_h(x), _h(w), b, 2, 1)))
    src = []

    def conv_dw(name, x, stride, r, r_out, packed):
        w, b = _fold_nb(sd, name + '.0', name + '.1')
        x = _dw_packed_half(x, w, b, stride) if packed else r(F.relu(F.conv2d(x, w, b, stride, 1, 1, x.size(1))))
        w, b = _fold_nb(sd, name + '.3', name + '.4')
        return r_out(F.relu(F.conv2d(x, r(w), b)))

    for n, (i, o, s) in enumerate(M.MOBILENET_DW):
        if n + 1 == 12:
            src.append(_bf(M.l2norm(x, sd['L2Norm_4_3.weight'])))
        x = conv_dw('backbone.%d' % (n + 1), x, s, _h, _h, True)
    src.append(_bf(M.l2norm(x, sd['L2Norm_5_3.weight'])))
    w, b = _fold_nb(sd, 'extras.0.0', 'extras.0.1')
    x = _h(F.relu(F.conv2d(x, _h(w), b)))
    x = conv_dw('extras.0.3', x, 2, _h, _bf, True)
    src.append(x)
    w, b = _fold_nb(sd, 'extras.1.0', 'extras.1.1')
    x = _bf(F.relu(F.conv2d(x, _bf(w), b)))
    x = conv_dw('extras.1.3', x, 2, _bf, _bf, False)
    src.append(x)
    return src


def test_mobilenet_half_trunk_control():
    """(1) The half trunk keeps the ARM regression within 1e-2 of the oracle (a bf16 trunk: 2.5e-2).  (2) The oracle's OWN heads, fed
    offsets regressed from that ARM regression, move by MORE than 2 x 2e-2 on rows whose taps did not change side: on this
    variant the reference's response to an 8e-3 offset perturbation is steep (measured here: odm_loc 5-10e-2, conf 1.1-1.5e-1),
    which is why the GPU gate of the MobileNet variant bounds the non-flipped rows by the size measured HERE and holds the
    product to 2e-2 on every row only against the oracle heads fed the product's own offsets (parity_tools.assert_bf16_gate (A))."""
    mod_name, spec_fn, build_kw, spec_kw, _ = CASES['drn_mobilenet320']
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    x = make_input(1, 320)
    with torch.no_grad():
        src = mobilenet_trunk_half_emulated(sd, x)
        arm = [F.conv2d(_bf(src[k]), _bf(sd['arm_loc.%d.weight' % k]), None, 1, 1) for k in range(4)]
        arm_flat = torch.cat([m.permute(0, 2, 3, 1).reshape(1, -1) for m in arm], 1).view(1, -1, 4)
        R = PT.drn_reference_bundle(sd, x, arm_flat, 21, False, [(s, s) for s in (40, 20, 10, 5)], trunk=M._mobilenet_trunk)
    e_arm = float((arm_flat - R['arm_loc']).abs().max() / R['arm_loc'].abs().max())
    assert e_arm < 1e-2, e_arm
    fl = R['flipped'].reshape(-1)
    rep_l = PT.split_report(R['odm_loc_given'].reshape(-1, 4).numpy(), R['odm_loc'].reshape(-1, 4).numpy(), fl, TOL)
    rep_c = PT.split_report(R['conf_given'].numpy(), R['conf'].numpy(), fl, TOL)
    assert 4e-2 < rep_l['max_other'] < MOBILE_SLACK * TOL and 4e-2 < rep_c['max_other'] < MOBILE_SLACK * TOL, (rep_l, rep_c)
    assert rep_l['n_beyond'] < 0.03 * rep_l['rows'] and rep_c['n_beyond'] < 0.03 * rep_c['rows'], (rep_l, rep_c)
    assert rep_l['l2'] < 3e-2 and rep_c['l2'] < 3e-2


MOBILE_SLACK = 10.0     # bound on the non-flipped rows of the MobileNet variant, in units of 2e-2 (the control above measures 2.8-7.6)
