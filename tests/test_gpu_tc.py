"""GPU parity: tcgen05 kernels (implicit-GEMM conv, fused deformable head) vs fp32 references computed
from the SAME bf16-rounded operands (so the only difference is fp32 accumulation order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize('b,cin,cout,k,pad,dil,h,w,relu', [
    (2, 64, 64, 3, 1, 1, 32, 32, True),        # conv1_2 class (BN=64 tile)
    (1, 64, 128, 3, 1, 1, 24, 40, True),       # conv2_1 class (BN=128)
    (2, 128, 256, 3, 1, 1, 20, 20, True),      # conv3 class (BN=256), 20x20 boxes
    (1, 256, 512, 3, 1, 1, 40, 40, True),      # conv4 class: two N tiles, 40-wide rows
    (3, 512, 1024, 3, 6, 6, 10, 10, True),     # conv6: dilation 6
    (3, 1024, 256, 1, 0, 1, 10, 10, True),     # extras.0 / conv7 class: 1x1
    (7, 512, 256, 3, 1, 1, 5, 5, False),       # last_layer_trans: 5x5 maps, several images per tile
    (2, 64, 64, 3, 1, 1, 17, 23, False),       # ragged sizes
    (1, 192, 96, 5, 2, 1, 9, 9, False),        # 5x5 kernel, Cout not a multiple of 64
    (4, 64, 64, 3, 1, 1, 96, 96, True),        # 288 tiles > #SMs: persistent loop, both TMEM buffers, ring wrap
    (8, 128, 512, 3, 1, 1, 40, 40, True),      # 224 (m,n) tiles, BN=256
    (5, 64, 128, 3, 1, 1, 48, 48, False),      # BN=128, 6-stage ring
    (2, 128, 32, 3, 1, 1, 32, 24, True),       # halo kernel: two resident channel blocks, Cout < 64
    (1, 64, 48, 3, 1, 1, 16, 8, False),        # halo kernel: a single 8x16 tile, Cout = 48 (three 16-column chunks)
    (3, 64, 128, 3, 1, 1, 64, 64, True),       # halo kernel BN=128 (conv2_1 class), 96 tiles
    (4, 32, 64, 1, 0, 1, 40, 40, True),        # Cin = 32 (MobileNet pw1): 64-channel TMA box zero-fills channels 32..63
    (2, 96, 64, 3, 1, 1, 16, 16, False),       # Cin = 96: second channel block half out of bounds
    (2, 128, 256, 3, 1, 1, 32, 48, True),      # streamed halo kernel: two channel blocks, two 128-wide N tiles (conv3_1 class)
    (5, 128, 128, 3, 1, 1, 64, 96, False),     # streamed halo kernel: 120 units, both TMEM buffers and the 7-stage weight ring wrap
    (9, 512, 320, 3, 1, 1, 40, 40, True),      # two-M-tile units (MT=2): 121 M tiles (odd: last unit half empty), ragged-tail tiles, N tiles 256 + 64
    (21, 512, 512, 3, 1, 1, 20, 20, False),    # MT=2 on 20x20 maps (conv5 class): 71 M tiles incl. tail tiles of 3 images
    (16, 256, 600, 1, 0, 1, 40, 40, False),    # resident-weight mode (4 k-blocks = 4 stages, 600 tiles): contiguous tile ranges, 3 N tiles, last one 96 wide
])
def test_conv_tc_vs_fp32(b, cin, cout, k, pad, dil, h, w, relu):
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + cout + k)
    x = _bf(torch.randn(b, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, wt, bias, 1, pad, dil)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, bias, None, 1, pad, dil, device='cuda')
    out = ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc, relu=relu, out_dtype=torch.float32, use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5
    out16 = ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc, relu=relu, use_tc=True)
    assert out16.dtype == torch.bfloat16
    assert rel_err(_nchw(out16.float()).cpu().numpy(), ref.numpy()) < 6e-3     # bf16 output rounding


@pytest.mark.parametrize('b,cin,cout,h,w', [(2, 64, 64, 32, 32), (3, 64, 128, 48, 80), (1, 256, 256, 80, 80), (2, 128, 128, 160, 160)])
def test_conv_tc_fused_maxpool(b, cin, cout, h, w):
    """conv + BN-folded bias + ReLU + MaxPool2d(2,2) in one kernel (vgg() 'M'/'C' after conv1_2/2_2/3_3)."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + h)
    x = _bf(torch.randn(b, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.max_pool2d(F.relu(F.conv2d(x, wt, bias, 1, 1)), 2, 2)
    pc = ops.PackedConv(wt, bias, None, 1, 1, 1, device='cuda')
    out = ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc, relu=True, out_dtype=torch.float32, use_tc=True, pool=True)
    assert tuple(out.shape) == (b, h // 2, w // 2, cout)
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5


@pytest.mark.parametrize('b,cin,cout,h,w', [(3, 256, 512, 10, 10), (2, 64, 96, 17, 23), (1, 128, 64, 40, 40)])
def test_conv_tc_stride2(b, cin, cout, h, w):
    """3x3 stride-2 pad-1 conv (extras.3) through TMA element strides."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + h + 1)
    x = _bf(torch.randn(b, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, wt, bias, 2, 1))
    pc = ops.PackedConv(wt, bias, None, 2, 1, 1, device='cuda')
    out = ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc, relu=True, out_dtype=torch.float32, use_tc=True)
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5


def test_conv_tc_head_into_flat_buffer_and_deconv():
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(9)
    # ARM head: Cout=12, fp32 rows written at a prior offset of a flat [B,P,4] tensor
    B, H, W, P, off = 2, 10, 10, 6375, 6000
    x = _bf(torch.randn(B, 128, H, W, generator=g))
    wt = _bf(torch.randn(12, 128, 3, 3, generator=g) * 0.05)
    bias = torch.randn(12, generator=g)
    ref = F.conv2d(x, wt, bias, 1, 1).permute(0, 2, 3, 1).reshape(B, -1)
    pc = ops.PackedConv(wt, bias, None, 1, 1, 1, device='cuda')
    flat = torch.zeros(B, P * 4, device='cuda')
    ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc, out=flat[:, off * 4:], out_sb=P * 4, out_sp=12, use_tc=True)
    assert rel_err(flat[:, off * 4:off * 4 + H * W * 12].cpu().numpy(), ref.numpy()) < 2e-5
    assert not flat[:, :off * 4].any() and not flat[:, off * 4 + H * W * 12:].any()
    # ConvTranspose2d k2 s2 + residual + ReLU
    xd = _bf(torch.randn(2, 256, 5, 5, generator=g))
    wd = _bf(torch.randn(256, 256, 2, 2, generator=g) * 0.06)
    bd = torch.randn(256, generator=g) * 0.1
    t = _bf(torch.randn(2, 256, 10, 10, generator=g))
    ref = F.relu(F.conv_transpose2d(xd, wd, bd, 2, 0) + t)
    pd = ops.PackedConv(wd, bd, None, deconv=True, device='cuda')
    out = ops.conv2d(_nhwc(xd).cuda().to(torch.bfloat16), pd, relu=True, residual=_nhwc(t).cuda().to(torch.bfloat16), use_tc=True)
    assert rel_err(_nchw(out.float()).cpu().numpy(), ref.numpy()) < 6e-3


@pytest.mark.parametrize('B,H,W,cin,C,dg,multihead', [
    (2, 10, 10, 256, 21, 1, False), (1, 40, 40, 256, 21, 1, True), (3, 5, 5, 256, 21, 1, True),
    (1, 16, 16, 256, 81, 1, False), (2, 10, 10, 512, 31, 8, False), (1, 10, 10, 1024, 31, 8, False)])
def test_deform_head_tc_vs_oracle(B, H, W, cin, C, dg, multihead):
    from oracle import deform_conv_ref as R
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(B * H + C)
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    wl = _bf(torch.randn(12, cin, 3, 3, generator=g) * 0.03)
    wc = _bf(torch.randn(3 * C, cin, 3, 3, generator=g) * 0.03)
    off = torch.randn(B, dg * 18, H, W, generator=g) * 1.5
    loc_ref = R.deform_conv_forward(x, off, wl, 1, 1, 1, dg)
    conf_ref = R.deform_conv_forward(x, off, wc, 1, 1, 1, dg)
    w2 = off2 = None
    if multihead:
        wl2 = _bf(torch.randn(12, cin, 5, 5, generator=g) * 0.02)
        wc2 = _bf(torch.randn(3 * C, cin, 5, 5, generator=g) * 0.02)
        off2 = torch.randn(B, dg * 50, H, W, generator=g) * 1.5
        loc_ref = loc_ref + R.deform_conv_forward(x, off2, wl2, 1, 2, 1, dg)
        conf_ref = conf_ref + R.deform_conv_forward(x, off2, wc2, 1, 2, 1, dg)
        w2 = torch.cat([wl2, wc2], 0)

    pack = ops.pack_deform_head_weight

    P, poff = H * W * 3 + 11, 5
    loc = torch.zeros(B, P, 4, device='cuda')
    conf = torch.zeros(B, P, C, device='cuda')
    ops.deform_head(_nhwc(x).cuda().to(torch.bfloat16), _nhwc(off).cuda(), pack(torch.cat([wl, wc], 0)), C, dg, 3, 1,
                    loc, conf, P, poff, offsets2=_nhwc(off2).cuda() if multihead else None,
                    w2_bf16=pack(w2) if multihead else None, kh2=5 if multihead else 0, pad2=2 if multihead else 0,
                    softmax=False)
    torch.cuda.synchronize()
    lr = loc_ref.permute(0, 2, 3, 1).reshape(B, H * W * 3, 4)
    cr = conf_ref.permute(0, 2, 3, 1).reshape(B, H * W * 3, C)
    # the sampled A tile is rounded to bf16 before the MMA -> ~2^-9 relative per element
    assert rel_err(loc[:, poff:poff + H * W * 3].cpu().numpy(), lr.numpy()) < 1e-2
    assert rel_err(conf[:, poff:poff + H * W * 3].cpu().numpy(), cr.numpy()) < 1e-2
    assert not loc[:, :poff].any() and not loc[:, poff + H * W * 3:].any()
    conf2 = torch.zeros(B, P, C, device='cuda')
    ops.deform_head(_nhwc(x).cuda().to(torch.bfloat16), _nhwc(off).cuda(), pack(torch.cat([wl, wc], 0)), C, dg, 3, 1,
                    loc, conf2, P, poff, offsets2=_nhwc(off2).cuda() if multihead else None,
                    w2_bf16=pack(w2) if multihead else None, kh2=5 if multihead else 0, pad2=2 if multihead else 0,
                    softmax=True)
    sm = torch.softmax(conf[:, poff:poff + H * W * 3], -1)
    assert rel_err(conf2[:, poff:poff + H * W * 3].cpu().numpy(), sm.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('B,H,W,cin,C,multihead,chunk_mb', [
    (2, 10, 10, 256, 21, False, 48), (1, 40, 40, 256, 21, True, 48), (3, 5, 5, 256, 21, True, 48),
    (5, 20, 20, 256, 21, True, 5),           # 2.2 MB of projections per image -> chunks of 2, 2, 1 images
    (2, 13, 7, 256, 6, True, 48), (1, 16, 16, 512, 31, False, 48)])
def test_deform_head_projected_vs_oracle(B, H, W, cin, C, multihead, chunk_mb, monkeypatch):
    """Project-then-sample head (tdrn_conv2d_tc 1x1 per-tap projections + tdrn_deform_head_sample) against the
    restated reference loop (deform_conv_cuda.c:157-193), and against the fused im2col head on the same inputs."""
    from oracle import deform_conv_ref as R
    from tdrn_b200 import ops
    monkeypatch.setenv('TDRN_DEFORM_CHUNK_MB', str(chunk_mb))
    g = torch.Generator().manual_seed(B * H + C + 1)
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    wl = _bf(torch.randn(12, cin, 3, 3, generator=g) * 0.03)
    wc = _bf(torch.randn(3 * C, cin, 3, 3, generator=g) * 0.03)
    off = torch.randn(B, 18, H, W, generator=g) * 1.5
    loc_ref = R.deform_conv_forward(x, off, wl, 1, 1, 1, 1)
    conf_ref = R.deform_conv_forward(x, off, wc, 1, 1, 1, 1)
    w1, w2, off2 = torch.cat([wl, wc], 0), None, None
    if multihead:
        wl2 = _bf(torch.randn(12, cin, 5, 5, generator=g) * 0.02)
        wc2 = _bf(torch.randn(3 * C, cin, 5, 5, generator=g) * 0.02)
        off2 = torch.randn(B, 50, H, W, generator=g) * 1.5
        loc_ref = loc_ref + R.deform_conv_forward(x, off2, wl2, 1, 2, 1, 1)
        conf_ref = conf_ref + R.deform_conv_forward(x, off2, wc2, 1, 2, 1, 1)
        w2 = torch.cat([wl2, wc2], 0)
    pc, n_pad = ops.pack_deform_proj_weight(w1, w2)
    assert n_pad % 8 == 0 and pc.cout == (9 + (25 if multihead else 0)) * n_pad
    P, poff = H * W * 3 + 11, 5
    xb = _nhwc(x).cuda().to(torch.bfloat16)
    kw = dict(offsets2=_nhwc(off2).cuda() if multihead else None, kh2=5 if multihead else 0, pad2=2 if multihead else 0)
    loc = torch.zeros(B, P, 4, device='cuda')
    conf = torch.zeros(B, P, C, device='cuda')
    ops.deform_head_projected(xb, _nhwc(off).cuda(), pc, n_pad, C, 3, 1, loc, conf, P, poff, softmax=False, **kw)
    torch.cuda.synchronize()
    lr = loc_ref.permute(0, 2, 3, 1).reshape(B, H * W * 3, 4)
    cr = conf_ref.permute(0, 2, 3, 1).reshape(B, H * W * 3, C)
    # the per-tap projections are rounded to bf16 before they are sampled -> ~2^-9 relative per term
    assert rel_err(loc[:, poff:poff + H * W * 3].cpu().numpy(), lr.numpy()) < 1e-2
    assert rel_err(conf[:, poff:poff + H * W * 3].cpu().numpy(), cr.numpy()) < 1e-2
    assert not loc[:, :poff].any() and not loc[:, poff + H * W * 3:].any()
    assert not conf[:, :poff].any() and not conf[:, poff + H * W * 3:].any()
    conf2 = torch.zeros(B, P, C, device='cuda')
    ops.deform_head_projected(xb, _nhwc(off).cuda(), pc, n_pad, C, 3, 1, loc, conf2, P, poff, softmax=True, **kw)
    sm = torch.softmax(conf[:, poff:poff + H * W * 3], -1)
    assert rel_err(conf2[:, poff:poff + H * W * 3].cpu().numpy(), sm.cpu().numpy()) < 1e-5
    # same head through the fused im2col kernel: two bf16 roundings of different quantities, both within tolerance
    loc3 = torch.zeros(B, P, 4, device='cuda')
    conf3 = torch.zeros(B, P, C, device='cuda')
    pack = ops.pack_deform_head_weight
    ops.deform_head(xb, _nhwc(off).cuda(), pack(w1), C, 1, 3, 1, loc3, conf3, P, poff,
                    w2_bf16=pack(w2) if multihead else None, softmax=False, **kw)
    assert rel_err(loc.cpu().numpy(), loc3.cpu().numpy()) < 1.5e-2
    assert rel_err(conf.cpu().numpy(), conf3.cpu().numpy()) < 1.5e-2


@pytest.mark.parametrize('multihead,split,softmax', [(False, False, True), (True, False, True), (True, True, False),
                                                     (False, True, True)])
def test_deform_sample_group_is_bit_identical_to_per_level(multihead, split, softmax):
    """tdrn_deform_head_sample_group (all pyramid levels in one launch) writes exactly what the per-level launches write."""
    from tdrn_b200 import ops
    B, C, cin = 3, 21, 256
    sizes = [(20, 24), (10, 12), (5, 6), (3, 3)]
    g = torch.Generator().manual_seed(7 + multihead + 2 * split)
    P = sum(h * w * 3 for h, w in sizes)
    lv = [0]
    for h, w in sizes:
        lv.append(lv[-1] + h * w * 3)
    feats, offs, offs2, packs = [], [], [], []
    for h, w in sizes:
        x = torch.randn(B, h, w, cin, generator=g).cuda()
        feats.append(x if split else x.to(torch.bfloat16))
        offs.append((torch.randn(B, h, w, 18, generator=g) * 1.5).cuda())
        offs2.append((torch.randn(B, h, w, 50, generator=g) * 1.5).cuda())
        w1 = torch.randn(12 + 3 * C, cin, 3, 3, generator=g) * 0.03
        w2 = torch.randn(12 + 3 * C, cin, 5, 5, generator=g) * 0.02 if multihead else None
        packs.append(ops.pack_deform_proj_weight(w1, w2, x3=split))
    kw = dict(kh2=5 if multihead else 0, pad2=2 if multihead else 0, softmax=softmax, split=split)
    loc_a = torch.zeros(B, P, 4, device='cuda')
    conf_a = torch.zeros(B, P, C, device='cuda')
    for k in range(len(sizes)):
        ops.deform_head_projected(feats[k], offs[k], packs[k][0], packs[k][1], C, 3, 1, loc_a, conf_a, P, lv[k],
                                  offsets2=offs2[k] if multihead else None, **kw)
    loc_b = torch.zeros(B, P, 4, device='cuda')
    conf_b = torch.zeros(B, P, C, device='cuda')
    ys = [ops.deform_project(feats[k], packs[k][0], packs[k][1], 3, 5 if multihead else 0, C, split=split)
          for k in range(len(sizes))]
    ops.deform_sample_group(ys, [tuple(f.shape) for f in feats], packs[0][1], C, 3, 1, offs, loc_b, conf_b, P, lv[:-1],
                            offsets2=offs2 if multihead else None, **kw)
    torch.cuda.synchronize()
    assert loc_a.abs().sum() > 0 and torch.equal(loc_a, loc_b) and torch.equal(conf_a, conf_b)


def test_deform_heads_with_half_projections(monkeypatch):
    """Opt-in TDRN_PROJ_F16=1: the per-tap projections are IEEE half (tdrn_conv2d_tc out_dtype TDRN_F16, tdrn_deform_head_desc.split = 2).
    Same heads as with bf16 projections to within the projections' rounding; grouped launch == per-level launches bit for bit."""
    from tdrn_b200 import ops
    B, C, cin = 2, 21, 256
    sizes = [(20, 24), (5, 6)]
    g = torch.Generator().manual_seed(11)
    P = sum(h * w * 3 for h, w in sizes)
    lv = [0, sizes[0][0] * sizes[0][1] * 3]
    feats = [torch.randn(B, h, w, cin, generator=g).to(torch.bfloat16).cuda() for h, w in sizes]
    offs = [(torch.randn(B, h, w, 18, generator=g) * 1.5).cuda() for h, w in sizes]
    packs = [ops.pack_deform_proj_weight(torch.randn(12 + 3 * C, cin, 3, 3, generator=g) * 0.03, None) for _ in sizes]
    res = {}
    for mode in ('0', '1'):
        monkeypatch.setenv('TDRN_PROJ_F16', mode)
        loc = torch.zeros(B, P, 4, device='cuda'); conf = torch.zeros(B, P, C, device='cuda')
        for k in range(2):
            ops.deform_head_projected(feats[k], offs[k], packs[k][0], packs[k][1], C, 3, 1, loc, conf, P, lv[k], softmax=False)
        ys = [ops.deform_project(feats[k], packs[k][0], packs[k][1], 3, 0, C) for k in range(2)]
        assert ys[0].dtype == (torch.float16 if mode == '1' else torch.bfloat16)
        loc_g = torch.zeros(B, P, 4, device='cuda'); conf_g = torch.zeros(B, P, C, device='cuda')
        ops.deform_sample_group(ys, [tuple(f.shape) for f in feats], packs[0][1], C, 3, 1, offs, loc_g, conf_g, P, lv, softmax=False)
        torch.cuda.synchronize()
        assert torch.equal(loc, loc_g) and torch.equal(conf, conf_g)
        res[mode] = (loc.cpu().numpy(), conf.cpu().numpy())
    assert rel_err(res['1'][0], res['0'][0]) < 4e-3 and rel_err(res['1'][1], res['0'][1]) < 4e-3
    assert not np.array_equal(res['1'][1], res['0'][1])


@pytest.mark.parametrize('b,h,w,relu', [
    (2, 8, 64, True),          # 64x2 tiles, one tile row per image pair
    (3, 20, 96, True),         # 32x4 tiles
    (1, 24, 48, False),        # 16x8 tiles, negative outputs kept
    (5, 64, 128, True),        # 640 tiles > 2 x #SMs: persistent loop wraps, both TMEM/staging buffers reused
    (1, 10, 20, True),         # does not tile -> CUDA-core stem (same result contract)
])
def test_conv_stem_tc_vs_fp32(b, h, w, relu):
    """conv1_1 as a K=32 tcgen05 GEMM (csrc/conv_stem_tc.cu): fp32 NCHW image in, NHWC bf16 out, bias+ReLU fused.
    Reference: fp32 conv of the bf16-rounded image and weights (the MMA operands); 4e-3 = bf16 output rounding."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b * 100 + h + w)
    x = torch.randn(b, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(64, generator=g)
    ref = F.conv2d(_bf(x), _bf(wt), bias, 1, 1)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, bias, None, 1, 1, 1, device='cuda', want_bf16=False)
    out = ops.conv_first(x.cuda(), pc, relu, torch.bfloat16)
    assert out.shape == (b, h, w, 64) and out.dtype == torch.bfloat16
    tol = 4e-3 if (w % 16 == 0 and h % 2 == 0) else 8e-3      # the CUDA-core fallback multiplies un-rounded fp32 operands
    assert rel_err(_nchw(out.float()).cpu().numpy(), ref.numpy()) < tol


@pytest.mark.parametrize('b,h,w,cin,cout,stride', [
    (2, 32, 32, 32, 64, 1),          # backbone[1]: Cin = 32 (half a channel block), double-buffered accumulator sets
    (2, 160, 160, 32, 64, 1),        # the same at full size: 32 x 4 tiles
    (3, 80, 80, 128, 256, 1),        # 16 x 8 tiles
    (2, 40, 40, 512, 512, 1),        # backbone[7..11]: two 256-column sub-tiles, 40 x 3 tiles (partial at the bottom)
    (2, 20, 20, 512, 1024, 1),       # Cout = 1024: two passes over the couts
    (3, 20, 20, 1024, 1024, 1),      # 16 k-blocks
    (9, 10, 10, 256, 512, 1),        # one image per tile, map smaller than the tile
    (9, 5, 5, 256, 512, 1),          # several images per tile
    (1, 7, 9, 64, 72, 1),            # ragged everything: Cout % 16 != 0, odd map
    (40, 16, 16, 128, 256, 1),       # more tiles than SMs: the rings and both accumulator sets wrap
])
def test_conv_dwpw_is_bit_identical_to_the_two_kernel_path(b, h, w, cin, cout, stride):
    """tdrn_conv_dwpw (conv_dw block of networks.py:736-745 in one kernel) == tdrn_dwconv3x3 followed by tdrn_conv2d_tc, bit
    for bit, and both close to the fp32 reference of the block."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b + h + cin + cout)
    x = _bf(torch.randn(b, cin, h, w, generator=g))
    wd = torch.randn(cin, 1, 3, 3, generator=g) * 0.4
    bn1 = [torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.1, torch.randn(cin, generator=g) * 0.1,
           torch.rand(cin, generator=g) + 0.5]
    wp = torch.randn(cout, cin, 1, 1, generator=g) * (2.0 / cin) ** 0.5
    bn2 = [torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1, torch.randn(cout, generator=g) * 0.1,
           torch.rand(cout, generator=g) + 0.5]
    pd = ops.PackedDw(wd, bn1, stride, 'cuda')
    pc = ops.PackedConv(wp, None, bn2, 1, 0, 1, device='cuda')
    xb = _nhwc(x).cuda().to(torch.bfloat16)
    mid = ops.dwconv3x3(xb, pd, relu=True)
    want = ops.conv2d(mid, pc, relu=True, use_tc=True)
    for _ in range(2):
        got = ops.conv_dwpw(xb, pd, pc)
        torch.cuda.synchronize()
        assert got.shape == want.shape and got.dtype == torch.bfloat16
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    ref = F.relu(F.batch_norm(F.conv2d(x, wd, None, stride, 1, 1, cin), bn1[2], bn1[3], bn1[0], bn1[1], False, 0.0, 1e-5))
    ref = F.relu(F.batch_norm(F.conv2d(ref, wp), bn2[2], bn2[3], bn2[0], bn2[1], False, 0.0, 1e-5))
    assert rel_err(_nchw(got.float()).cpu().numpy(), ref.numpy()) < 2e-2


def test_conv_dwpw_refuses_stride_2():
    from tdrn_b200 import ops, _lib
    g = torch.Generator().manual_seed(3)
    bn = lambda c: [torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]
    pd = ops.PackedDw(torch.randn(64, 1, 3, 3, generator=g), bn(64), 2, 'cuda')
    pc = ops.PackedConv(torch.randn(128, 64, 1, 1, generator=g), None, bn(128), 1, 0, 1, device='cuda')
    with pytest.raises(_lib.TdrnError):
        ops.conv_dwpw(torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16, device='cuda'), pd, pc)


@pytest.mark.parametrize('b,h,w,relu', [
    (2, 32, 128, True),        # 16 x 64 output: 64x2 tiles
    (3, 40, 192, True),        # 20 x 96 output: 32x4 tiles
    (1, 48, 96, False),        # 24 x 48 output: 16x8 tiles, negative outputs kept
    (4, 320, 320, True),       # the MobileNet-320 stem itself: 1 000 tiles per image batch > 2 x #SMs
    (1, 20, 40, True),         # does not tile -> CUDA-core stem (same result contract)
])
def test_conv_stem_tc_stride2_mobilenet(b, h, w, relu):
    """The MobileNet stem conv_bn(3, 32, 2) (dualrefinedet_mobilenet.py:20) on the tensor cores (csrc/conv_stem_tc.cu, S = 2,
    COUT = 32): fp32 NCHW image in, NHWC bf16 out.  Reference: fp32 conv of the bf16-rounded image and weights."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b * 100 + h + w + 2)
    x = torch.randn(b, 3, h, w, generator=g)
    wt = torch.randn(32, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(32, generator=g)
    ref = F.conv2d(_bf(x), _bf(wt), bias, 2, 1)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, bias, None, 2, 1, 1, device='cuda', want_bf16=False)
    out = ops.conv_first(x.cuda(), pc, relu, torch.bfloat16)
    torch.cuda.synchronize()
    assert out.shape == (b, h // 2, w // 2, 32) and out.dtype == torch.bfloat16
    tiles = (w // 2) % 16 == 0 and (h // 2) % 2 == 0
    assert rel_err(_nchw(out.float()).cpu().numpy(), ref.numpy()) < (4e-3 if tiles else 8e-3)


@pytest.mark.parametrize('b,h,w,relu', [(2, 16, 64, True), (3, 20, 96, True), (1, 24, 48, False), (5, 64, 128, True)])
def test_conv_stem_split_precision(b, h, w, relu):
    """conv1_1 on the tensor cores in split precision (out_dtype TDRN_BF16_SPLIT): fp32-accurate result, written as the
    (hi | lo) operand of conv1_2 -- against a float64 conv, and bit-identical in format to tdrn_split_bf16 of its own value."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b * 100 + h + w + 5)
    x = torch.randn(b, 3, h, w, generator=g) * 60.0                  # mean-subtracted pixel range
    wt = torch.randn(64, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(64, generator=g)
    ref = F.conv2d(x.double(), wt.double(), bias.double(), 1, 1)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, bias, None, 1, 1, 1, device='cuda', want_bf16=False)
    assert ops.conv_first_split_ok(x.cuda(), pc)
    out = ops.conv_first(x.cuda(), pc, relu, torch.float32, split=True)
    torch.cuda.synchronize()
    assert out.shape == (b, h, w, 128) and out.dtype == torch.bfloat16
    val = out[..., :64].float() + out[..., 64:].float()
    assert rel_err(_nchw(val).cpu().numpy(), ref.numpy()) < 1e-5
    # hi is the bf16 nearest to the value (re-splitting hi + lo gives hi back, except where lo rounded up to exactly half an ulp)
    again = ops.split_bf16(val.contiguous())
    assert float((again[..., :64].view(torch.int16) == out[..., :64].view(torch.int16)).float().mean()) > 0.99
    assert bool(((out[..., :64].float() - val).abs() <= val.abs() * 2.0 ** -8).all())
    # plain bf16 operands are two orders of magnitude further away
    out16 = ops.conv_first(x.cuda(), pc, relu, torch.bfloat16)
    assert rel_err(_nchw(out16.float()).cpu().numpy(), ref.numpy()) > 1e-3


@pytest.mark.parametrize('b,h,w,pool', [(2, 32, 32, True), (1, 16, 16, True), (3, 64, 48, False), (2, 48, 80, True),
                                        (5, 320, 320, True)])    # 5 x 800 tiles: every CTA runs many tiles, all rings wrap
def test_conv_stem_pair_is_bit_identical_to_the_two_kernel_path(b, h, w, pool):
    """conv1_1 + conv1_2 (+pool) fused (tdrn_conv_stem_pair) against conv_stem_tc -> conv_halo_kernel and against fp32."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b * h + w)
    x = torch.randn(b, 3, h, w, generator=g)
    w1 = torch.randn(64, 3, 3, 3, generator=g) * 0.2
    b1 = torch.randn(64, generator=g) * 0.1
    w2 = _bf(torch.randn(64, 64, 3, 3, generator=g) * 0.05)
    b2 = torch.randn(64, generator=g) * 0.1
    pc1 = ops.PackedConv(w1, b1, None, 1, 1, 1, device='cuda')
    pc2 = ops.PackedConv(w2, b2, None, 1, 1, 1, device='cuda')
    xc = x.cuda()
    y1 = ops.conv_first(xc, pc1, True, torch.bfloat16)
    ref2 = ops.conv2d(y1, pc2, relu=True, use_tc=True, pool=pool)
    got = ops.conv_stem_pair(xc, pc1, pc2, relu=True, pool=pool)
    torch.cuda.synchronize()
    assert got is not None and got.shape == ref2.shape and got.dtype == torch.bfloat16
    assert torch.equal(got, ref2)
    # and against an fp32 reference computed from the same bf16-rounded operands of each stage
    a1 = _bf(F.relu(F.conv2d(_bf(x), _bf(w1), b1, 1, 1)))
    a2 = F.relu(F.conv2d(a1, w2, b2, 1, 1))
    if pool:
        a2 = F.max_pool2d(a2, 2, 2)
    assert rel_err(_nchw(got.float()).cpu().numpy(), a2.numpy()) < 2e-2
    assert ops.conv_stem_pair(torch.randn(1, 3, 20, 12).cuda(), pc1, pc2) is None      # does not tile: caller falls back


# ---------------------------------------------------------------------------------------------------------------
# fp32-accurate tensor-core path (tdrn_conv_desc.split3): x = hi + lo, w = hi + lo (bf16 each),
# hi*W_hi + hi*W_lo + lo*W_hi accumulated in fp32 on tcgen05.  Inputs here are full fp32 (NOT pre-rounded).
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('b,cin,cout,k,stride,pad,dil,h,w,relu,pool', [
    (2, 64, 64, 3, 1, 1, 1, 32, 32, True, False),        # conv1_2 class
    (2, 64, 64, 3, 1, 1, 1, 32, 32, True, True),         # + fused 2x2 max-pool, fp32 output
    (1, 256, 512, 3, 1, 1, 1, 40, 40, True, False),      # conv4 class, ragged-tail tiles, 36 x 3 k-blocks
    (3, 512, 1024, 3, 1, 6, 6, 10, 10, True, False),     # conv6: dilation 6
    (3, 1024, 256, 1, 1, 0, 1, 10, 10, True, False),     # 1x1
    (3, 256, 512, 3, 2, 1, 1, 10, 10, True, False),      # extras.3: stride 2
    (2, 512, 12, 3, 1, 1, 1, 20, 20, False, False),      # ARM head: Cout = 12
    (16, 256, 600, 1, 1, 0, 1, 40, 40, False, False),    # many tiles, several N tiles
])
def test_conv_tc_split3_fp32_accuracy(b, cin, cout, k, stride, pad, dil, h, w, relu, pool):
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + cout + k + h)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x.double(), wt.double(), bias.double(), stride, pad, dil)
    if relu:
        ref = F.relu(ref)
    if pool:
        ref = F.max_pool2d(ref, 2, 2)
    pc = ops.PackedConv(wt, bias, None, stride, pad, dil, device='cuda', want_bf16=False, want_x3=True)
    xs = ops.split_bf16(_nhwc(x).cuda())
    hi, lo = xs[..., :cin].float(), xs[..., cin:].float()
    assert torch.equal(hi, _nhwc(x).cuda().to(torch.bfloat16).float())
    assert float(((hi + lo) - _nhwc(x).cuda()).abs().max()) <= 2.0 ** -16 * float(x.abs().max())   # 16 mantissa bits
    out = ops.conv2d(xs, pc, relu=relu, out_dtype=torch.float32, use_tc=True, split3=True, pool=pool)
    torch.cuda.synchronize()
    assert out.dtype == torch.float32
    # a single fp32 conv on cuDNN / oneDNN is itself ~1e-6 from the float64 result; the split costs ~1e-5
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5
    # same layer through plain bf16 operands: two orders of magnitude further away (the test would notice a missing lo term)
    pc16 = ops.PackedConv(wt, bias, None, stride, pad, dil, device='cuda')
    out16 = ops.conv2d(_nhwc(x).cuda().to(torch.bfloat16), pc16, relu=relu, out_dtype=torch.float32, use_tc=True, pool=pool)
    assert rel_err(_nchw(out16).cpu().numpy(), ref.numpy()) > 2e-4


@pytest.mark.parametrize('b,cin,cout,k,stride,pad,dil,h,w,relu,pool', [
    (2, 64, 64, 3, 1, 1, 1, 32, 32, True, False),        # one 64-column group, double-buffered accumulator sets
    (3, 64, 64, 3, 1, 1, 1, 32, 48, True, True),         # + fused pool: 32-row boxes
    (2, 64, 128, 3, 1, 1, 1, 32, 32, True, False),       # two groups per tile
    (2, 128, 128, 3, 1, 1, 1, 16, 32, True, True),
    (1, 256, 512, 3, 1, 1, 1, 40, 40, True, False),      # ragged-tail tiles, four N tiles
    (3, 512, 1024, 3, 1, 6, 6, 10, 10, True, False),     # conv6: several images per tile
    (3, 256, 512, 3, 2, 1, 1, 10, 10, False, False),     # stride 2
    (2, 256, 80, 1, 1, 0, 1, 20, 20, False, False),      # Cout % 64 != 0: register-store path of the same output format
])
def test_conv_tc_split3_writes_the_next_layers_operand(b, cin, cout, k, stride, pad, dil, h, w, relu, pool):
    """split_out = Cout: the layer writes the (hi | lo) operand of the next fp32-accurate conv itself -- bit-identical to its
    fp32 output passed through tdrn_split_bf16."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + cout + k + h + 1)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    pc = ops.PackedConv(wt, bias, None, stride, pad, dil, device='cuda', want_bf16=False, want_x3=True)
    xs = ops.split_bf16(_nhwc(x).cuda())
    want = ops.split_bf16(ops.conv2d(xs, pc, relu=relu, out_dtype=torch.float32, use_tc=True, split3=True, pool=pool))
    for _ in range(2):
        got = ops.conv2d(xs, pc, relu=relu, use_tc=True, split3=True, pool=pool, split_out=cout)
        torch.cuda.synchronize()
        assert got.dtype == torch.bfloat16 and got.shape == want.shape
        assert torch.equal(got.view(torch.int16), want.view(torch.int16))


def test_conv_tc_split3_deconv_residual():
    """ConvTranspose2d k2 s2 + residual + ReLU (`relu(up(x) + t)`, dualrefinedet_vggbn.py:177) on the fp32-accurate path."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(77)
    B, C, H, W = 3, 256, 10, 10
    x = torch.randn(B, C, H, W, generator=g)
    wt = torch.randn(C, C, 2, 2, generator=g) * 0.05
    bias = torch.randn(C, generator=g) * 0.1
    t = torch.randn(B, C, 2 * H, 2 * W, generator=g)
    ref = F.relu(F.conv_transpose2d(x.double(), wt.double(), bias.double(), 2, 0) + t.double())
    pc = ops.PackedConv(wt, bias, None, deconv=True, device='cuda', want_bf16=False, want_x3=True)
    out = ops.conv2d(ops.split_bf16(_nhwc(x).cuda()), pc, relu=True, residual=_nhwc(t).cuda(), out_dtype=torch.float32,
                     use_tc=True, split3=True)
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5


@pytest.mark.parametrize('b,cin,cout,k,stride,pad,h,w,relu', [
    (8, 1024, 12, 3, 1, 1, 10, 10, False),      # arm_loc.2: 144 k-blocks over a cluster of 8, one N tile
    (8, 1024, 256, 3, 1, 1, 10, 10, True),      # trans_layers.2.0: cluster of 4 x 4 N tiles
    (9, 512, 256, 3, 1, 1, 5, 5, True),         # last_layer_trans.0: 5 images per tile, cluster of 4
    (2, 256, 256, 3, 1, 1, 16, 16, True),       # 16x16 maps (512-pixel input): two tiles per image
    (3, 256, 512, 3, 2, 1, 10, 10, True),       # extras.3: stride 2, cluster of 2
    (1, 1024, 256, 1, 1, 0, 10, 10, True),      # a single tile (batch 1): 16 k-blocks over 2 CTAs
])
def test_conv_splitk_cluster(monkeypatch, b, cin, cout, k, stride, pad, h, w, relu):
    """Small maps: K split over a thread-block cluster, partial tiles summed through distributed shared memory
    (conv_splitk_kernel).  Same reference and tolerance as the plain kernel; the result of an image must not depend on
    the batch it is in (the cluster size is a function of the layer's shape only)."""
    from tdrn_b200 import ops, _lib
    g = torch.Generator().manual_seed(cin + cout + k + b)
    x = _bf(torch.randn(b, cin, h, w, generator=g))
    wt = _bf(torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, wt, bias, stride, pad)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, bias, None, stride, pad, 1, device='cuda')
    xg = _nhwc(x).cuda().to(torch.bfloat16)
    monkeypatch.setenv('TDRN_SPLITK', '1')                   # opt-in path (csrc/conv_tc.cu says why it is not the default)
    out = ops.conv2d(xg, pc, relu=relu, out_dtype=torch.float32, use_tc=True)
    torch.cuda.synchronize()
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 2e-5
    out16 = ops.conv2d(xg, pc, relu=relu, use_tc=True)
    assert rel_err(_nchw(out16.float()).cpu().numpy(), ref.numpy()) < 6e-3
    one = ops.conv2d(xg[b - 1:b].contiguous(), pc, relu=relu, out_dtype=torch.float32, use_tc=True)
    assert torch.equal(one[0], out[b - 1])
    monkeypatch.delenv('TDRN_SPLITK')
    plain = ops.conv2d(xg, pc, relu=relu, out_dtype=torch.float32, use_tc=True)      # the default kernel: same values to fp32 rounding
    assert rel_err(out.cpu().numpy(), plain.cpu().numpy()) < 1e-5
