"""GPU parity: fp32 SIMT operators called through the C ABI vs the oracle / plain torch fp32."""
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


@pytest.mark.parametrize('k,pad,dg,stride,dil,cin,cout,h,w', [
    (3, 1, 1, 1, 1, 8, 6, 9, 11), (3, 1, 2, 1, 1, 8, 12, 9, 11), (5, 2, 1, 1, 1, 8, 6, 7, 7),
    (3, 1, 4, 2, 1, 8, 6, 10, 9), (3, 2, 1, 1, 2, 4, 5, 9, 9), (3, 1, 8, 1, 1, 64, 75, 10, 10),
    (1, 0, 1, 1, 1, 6, 4, 5, 5)])
def test_deform_conv_forward_nchw_vs_oracle(k, pad, dg, stride, dil, cin, cout, h, w):
    """tdrn_deform_conv_forward == deform_conv_forward_cuda semantics (border rules included)."""
    from oracle import deform_conv_ref as R
    from tdrn_b200.model.networks import ConvOffset2d
    g = torch.Generator().manual_seed(k * 100 + dg)
    x = torch.randn(2, cin, h, w, generator=g)
    m = ConvOffset2d(cin, cout, k, stride=stride, padding=pad, dilation=dil, num_deformable_groups=dg)
    ho, wo = R.output_size(h, w, k, k, stride, pad, dil)
    off = torch.randn(2, dg * 2 * k * k, ho, wo, generator=g) * 2.5
    ref = R.deform_conv_forward(x, off, m.weight.detach(), stride, pad, dil, dg)
    out = m.cuda()(x.cuda(), off.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5          # fp32: accumulation order only


def test_deform_conv_border_probes():
    from tdrn_b200.model.networks import conv_offset2d
    x = (torch.arange(16, dtype=torch.float32).view(1, 1, 4, 4) + 1).cuda()
    w = torch.ones(1, 1, 1, 1).cuda()
    def at(dy, dx):
        off = torch.zeros(1, 2, 4, 4); off[0, 0] = dy; off[0, 1] = dx
        return conv_offset2d(x, off.cuda(), w)[0, 0].cpu()
    assert at(-0.5, 0.0)[0, 0] == 0
    assert at(0.5, 0.0)[3, 2] == 15
    assert at(0.0, 0.75)[1, 3] == 8
    assert at(1.0, 0.0)[3, 0] == 0


def test_deform_conv_shape_errors():
    from tdrn_b200.model.networks import conv_offset2d
    x = torch.zeros(1, 4, 5, 5).cuda()
    w = torch.zeros(3, 4, 3, 3).cuda()
    with pytest.raises(ValueError):
        conv_offset2d(torch.zeros(4, 5, 5).cuda(), torch.zeros(1, 18, 5, 5).cuda(), w, padding=1)
    with pytest.raises(RuntimeError):
        conv_offset2d(x, torch.zeros(1, 18, 4, 4).cuda(), w, padding=1)       # wrong offset map size
    with pytest.raises(ValueError):
        conv_offset2d(torch.zeros(1, 4, 2, 2).cuda(), torch.zeros(1, 18, 1, 1).cuda(), w)   # too small


@pytest.mark.parametrize('cin,cout,k,stride,pad,dil,h,w,relu,bias', [
    (3, 16, 3, 1, 1, 1, 13, 9, True, True), (16, 24, 3, 1, 1, 1, 12, 12, True, True),
    (32, 12, 3, 1, 1, 1, 10, 10, False, True), (16, 40, 1, 1, 0, 1, 7, 5, True, False),
    (16, 32, 3, 2, 1, 1, 10, 10, True, True), (16, 32, 3, 1, 6, 6, 10, 10, True, True),
    (12, 18, 1, 1, 0, 1, 5, 5, False, True), (8, 75, 5, 1, 2, 1, 6, 6, False, False)])
def test_conv2d_simt_vs_torch(cin, cout, k, stride, pad, dil, h, w, relu, bias):
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(3, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * 0.2
    b = torch.randn(cout, generator=g) if bias else None
    ref = F.conv2d(x, wt, b, stride, pad, dil)
    if relu:
        ref = F.relu(ref)
    pc = ops.PackedConv(wt, b, None, stride, pad, dil, device='cuda', want_bf16=False)
    out = ops.conv2d(_nhwc(x).cuda(), pc, relu=relu)
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 1e-5


def test_conv2d_bn_fold_residual_and_deconv():
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 6, 7, generator=g)
    wt = torch.randn(8, 16, 3, 3, generator=g) * 0.2
    b = torch.randn(8, generator=g)
    bn = (torch.rand(8, generator=g) + 0.5, torch.randn(8, generator=g), torch.randn(8, generator=g), torch.rand(8, generator=g) + 0.5)
    ref = F.relu(F.batch_norm(F.conv2d(x, wt, b, 1, 1), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
    pc = ops.PackedConv(wt, b, bn, 1, 1, 1, device='cuda', want_bf16=False)
    out = ops.conv2d(_nhwc(x).cuda(), pc, relu=True)
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 1e-5
    # ConvTranspose2d k2 s2 + residual + relu  (dualrefinedet_vggbn.py:177)
    wd = torch.randn(16, 8, 2, 2, generator=g) * 0.3
    bd = torch.randn(8, generator=g)
    t = torch.randn(2, 8, 12, 14, generator=g)
    ref = F.relu(F.conv_transpose2d(x, wd, bd, 2, 0) + t)
    pd = ops.PackedConv(wd, bd, None, deconv=True, device='cuda', want_bf16=False)
    out = ops.conv2d(_nhwc(x).cuda(), pd, relu=True, residual=_nhwc(t).cuda())
    assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 1e-5


def test_conv_first_nchw_input():
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 20, 24, generator=g)
    for stride, cout in ((1, 64), (2, 32)):
        wt = torch.randn(cout, 3, 3, 3, generator=g) * 0.3
        b = torch.randn(cout, generator=g)
        ref = F.relu(F.conv2d(x, wt, b, stride, 1))
        pc = ops.PackedConv(wt, b, None, stride, 1, 1, device='cuda', want_bf16=False)
        out = ops.conv_first(x.cuda(), pc, True, torch.float32)
        assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 1e-5


def test_pool_l2norm_softmax_dw_layout(golden):
    from tdrn_b200 import ops
    from tdrn_b200.layers.modules import L2Norm
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 24, 11, 13, generator=g)
    for ceil in (False, True):
        ref = F.max_pool2d(x, 2, 2, ceil_mode=ceil)
        out = ops.maxpool2x2(_nhwc(x).cuda(), ceil)
        assert torch.equal(_nchw(out).cpu(), ref)
    gs = golden('small_cases')
    l2 = L2Norm(16, 10)
    with torch.no_grad():
        l2.weight.copy_(torch.from_numpy(gs['l2_w']))
    y = l2.cuda()(torch.from_numpy(gs['l2_x']).cuda())
    assert rel_err(y.cpu().numpy(), gs['l2_y']) < 1e-6
    z = torch.randn(1000, 21, generator=g) * 3
    assert rel_err(ops.softmax_rows(z.cuda()).cpu().numpy(), F.softmax(z, 1).numpy()) < 1e-6
    wd = torch.randn(24, 1, 3, 3, generator=g)
    bn = (torch.rand(24, generator=g) + 0.5, torch.randn(24, generator=g), torch.randn(24, generator=g), torch.rand(24, generator=g) + 0.5)
    for stride in (1, 2):
        ref = F.relu(F.batch_norm(F.conv2d(x, wd, None, stride, 1, 1, 24), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
        out = ops.dwconv3x3(_nhwc(x).cuda(), ops.PackedDw(wd, bn, stride, 'cuda'))
        assert rel_err(_nchw(out).cpu().numpy(), ref.numpy()) < 1e-5
    assert torch.equal(ops.nhwc_to_nchw_f32(ops.nchw_f32_to_nhwc(x.cuda(), torch.float32)).cpu(), x)


@pytest.mark.parametrize('b,c,h,w', [(2, 512, 40, 40), (3, 256, 6, 10), (1, 1024, 20, 20), (2, 512, 5, 5)])
def test_l2norm_pool_fused(b, c, h, w):
    """tdrn_l2norm_pool2x2 == (L2Norm, MaxPool2d(2,2)) of the same bf16 input; odd sizes fall back to the two kernels."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(b, h, w, c, generator=g) * 3).to(torch.bfloat16)
    wt = torch.rand(c, generator=g) * 10 + 1
    xf = x.float()
    norm = xf.pow(2).sum(3, keepdim=True).sqrt() + 1e-10
    ref_n = (wt * (xf / norm)).to(torch.bfloat16)
    ref_p = F.max_pool2d(xf.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).to(torch.bfloat16)
    out_n, out_p = ops.l2norm_pool(x.cuda(), wt.cuda())
    assert torch.equal(out_p.cpu(), ref_p)
    assert rel_err(out_n.float().cpu().numpy(), ref_n.float().numpy()) < 8e-3          # one bf16 ulp


@pytest.mark.parametrize('stride', [1, 2])
@pytest.mark.parametrize('b,c,h,w', [(2, 24, 11, 13), (3, 64, 16, 16), (1, 256, 7, 9)])
def test_dwconv3x3_bf16_vectorised(b, c, h, w, stride):
    """Depthwise 3x3 + folded BN + ReLU (conv_dw first half, model/networks.py:738-740), bf16 fast path: 8 channels per
    thread, 4 output pixels per thread; fp32 reference on the bf16-rounded input; tolerance = bf16 output rounding."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(c + h + stride)
    x = torch.randn(b, c, h, w, generator=g).to(torch.bfloat16)
    wd = torch.randn(c, 1, 3, 3, generator=g)
    bn = (torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5)
    ref = F.relu(F.batch_norm(F.conv2d(x.float(), wd, None, stride, 1, 1, c), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
    out = ops.dwconv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.PackedDw(wd, bn, stride, 'cuda'))
    assert out.dtype == torch.bfloat16
    assert rel_err(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < 6e-3


def test_l2norm_bf16_vectorised():
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(77)
    for c, px in ((512, 40), (1024, 12), (256, 6)):
        x = (torch.randn(2, px, 2, c, generator=g) * 2).to(torch.bfloat16)
        wt = torch.rand(c, generator=g) * 20 + 1
        xf = x.float()
        ref = (wt * (xf / (xf.pow(2).sum(3, keepdim=True).sqrt() + 1e-10))).to(torch.bfloat16)
        out = ops.l2norm(x.cuda(), wt.cuda())
        assert rel_err(out.float().cpu().numpy(), ref.float().numpy()) < 8e-3


# ---------------------------------------------------------------------------------------------------------------
# IEEE-half variants of the MobileNet trunk operators (TDRN_F16): stem, depthwise 3x3, pointwise 1x1, L2Norm.
# References: fp32 torch on the half-rounded operands; tolerance = the output rounding of the output format.
# ---------------------------------------------------------------------------------------------------------------
def _h(x):
    return x.to(torch.float16).float()


@pytest.mark.parametrize('stride', [1, 2])
@pytest.mark.parametrize('din,dout', [(torch.float16, torch.float16), (torch.bfloat16, torch.float16), (torch.float16, torch.bfloat16)])
def test_dwconv3x3_half_formats(stride, din, dout):
    from tdrn_b200 import ops
    b, c, h, w = 2, 64, 18, 22
    g = torch.Generator().manual_seed(c + h + stride)
    x = torch.randn(b, c, h, w, generator=g).to(din)
    wd = torch.randn(c, 1, 3, 3, generator=g)
    bn = (torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5)
    ref = F.relu(F.batch_norm(F.conv2d(x.float(), wd, None, stride, 1, 1, c), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5))
    out = ops.dwconv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.PackedDw(wd, bn, stride, 'cuda'), out_dtype=dout)
    assert out.dtype == dout
    # half -> half runs the packed-half kernel (nine FMAs rounded to half each): a few half ulps; the mixed forms accumulate in fp32
    tol = 3e-3 if (din, dout) == (torch.float16, torch.float16) else (8e-4 if dout == torch.float16 else 6e-3)
    assert rel_err(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < tol
    # same input format in and out == the plain entry point
    if din == dout:
        assert torch.equal(out, ops.dwconv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), ops.PackedDw(wd, bn, stride, 'cuda')))


def test_dwconv3x3_half_saturates():
    """Conversions to half saturate at the largest finite value instead of producing Inf."""
    from tdrn_b200 import ops
    c = 8
    x = torch.full((1, 4, 4, c), 3.0e4, dtype=torch.float16)
    wd = torch.ones(c, 1, 3, 3)
    bn = (torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c))
    out = ops.dwconv3x3(x.cuda(), ops.PackedDw(wd, bn, 1, 'cuda'))
    assert bool(torch.isfinite(out).all()) and float(out.max()) == 65504.0


@pytest.mark.parametrize('b,cin,cout,h,w,dout', [(2, 64, 128, 20, 20, torch.float16), (3, 32, 64, 16, 24, torch.float16),
                                                 (2, 256, 512, 10, 10, torch.bfloat16), (1, 512, 512, 40, 40, torch.float16)])
def test_conv1x1_tc_half_operands(b, cin, cout, h, w, dout):
    """Pointwise conv of the half-precision trunk: half activations and half-packed weights on tcgen05 (kind::f16 with the
    half format codes), fp32 accumulation, half or bf16 output."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w, generator=g).to(torch.float16)
    wt = torch.randn(cout, cin, 1, 1, generator=g) * (cin ** -0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.float(), _h(wt), bias))
    pc = ops.PackedConv(wt, bias, None, 1, 0, 1, device='cuda', want_f16=True)
    out = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), pc, relu=True, use_tc=True, out_dtype=dout)
    torch.cuda.synchronize()
    assert out.dtype == dout and out.shape == (b, h, w, cout)
    assert rel_err(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < (8e-4 if dout == torch.float16 else 6e-3)
    # weights packed lazily from the fp32 copy give the same half bits
    pc2 = ops.PackedConv(wt, bias, None, 1, 0, 1, device='cuda')
    assert pc2.w_f16 is None and torch.equal(ops.ensure_f16(pc2), pc.w_f16)


@pytest.mark.parametrize('b,h,w', [(2, 32, 128), (1, 320, 320)])
def test_conv_stem_half_output(b, h, w):
    """MobileNet stem with half operands and output (tdrn_conv_first, out_dtype TDRN_F16)."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b + h)
    x = torch.randn(b, 3, h, w, generator=g) * 40.0          # pixel-scale inputs
    wt = torch.randn(32, 3, 3, 3, generator=g) * 0.05
    bias = torch.randn(32, generator=g)
    ref = F.relu(F.conv2d(_h(x), _h(wt), bias, 2, 1))
    pc = ops.PackedConv(wt, bias, None, 2, 1, 1, device='cuda', want_bf16=False)
    assert ops.conv_first_f16_ok(x.cuda())
    out = ops.conv_first(x.cuda(), pc, True, torch.float16)
    torch.cuda.synchronize()
    assert out.dtype == torch.float16 and out.shape == (b, h // 2, w // 2, 32)
    assert rel_err(out.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < 8e-4
    assert not ops.conv_first_f16_ok(torch.zeros(1, 3, 20, 40, device='cuda'))


def test_l2norm_half_in_bf16_out():
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(78)
    for c, px in ((512, 40), (1024, 12)):
        x = (torch.randn(2, px, 2, c, generator=g) * 2).to(torch.float16)
        wt = torch.rand(c, generator=g) * 20 + 1
        xf = x.float()
        ref = wt * (xf / (xf.pow(2).sum(3, keepdim=True).sqrt() + 1e-10))
        out = ops.l2norm(x.cuda(), wt.cuda(), out_dtype=torch.bfloat16)
        assert out.dtype == torch.bfloat16
        assert rel_err(out.float().cpu().numpy(), ref.numpy()) < 6e-3
        out16 = ops.l2norm(x.cuda(), wt.cuda())
        assert out16.dtype == torch.float16 and rel_err(out16.float().cpu().numpy(), ref.numpy()) < 8e-4


@pytest.mark.parametrize('b,c,h,w,relu', [(2, 64, 18, 22, True), (1, 256, 40, 40, True), (3, 32, 7, 5, False), (2, 1024, 20, 20, True), (1, 8, 9, 1, True)])
def test_dwconv3x3_half_row_walking_form_is_bit_identical(b, c, h, w, relu, monkeypatch):
    """The row-walking packed-half depthwise kernel (large batches; forced here with TDRN_DW_ROLL_MIN=1) performs the same FMAs in
    the same order as the one-row form: identical bits, on ragged maps too (H not a multiple of 8, W not a multiple of 4, W = 1)."""
    from tdrn_b200 import ops
    g = torch.Generator().manual_seed(b * 1000 + c + h)
    x = (torch.randn(b, h, w, c, generator=g) * 3).to(torch.float16).cuda()
    wd = torch.randn(c, 1, 3, 3, generator=g)
    bn = (torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5)
    pd = ops.PackedDw(wd, bn, 1, 'cuda')
    monkeypatch.setenv('TDRN_DW_ROLL_MIN', '0')
    one_row = ops.dwconv3x3(x, pd, relu=relu)
    monkeypatch.setenv('TDRN_DW_ROLL_MIN', '1')
    walking = ops.dwconv3x3(x, pd, relu=relu)
    torch.cuda.synchronize()
    assert torch.equal(one_row, walking)
    ref = F.batch_norm(F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), wd, None, 1, 1, 1, c), bn[2], bn[3], bn[0], bn[1], False, 0.0, 1e-5)
    if relu:
        ref = F.relu(ref)
    assert rel_err(walking.float().permute(0, 3, 1, 2).cpu().numpy(), ref.numpy()) < 3e-3
