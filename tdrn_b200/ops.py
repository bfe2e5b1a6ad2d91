"""Thin torch-tensor wrappers over the C ABI (include/tdrn_b200.h).

PyTorch supplies device memory and the current stream; every computation is a call into
libtdrn_b200.so.  Tensors named ``*_nhwc`` are [B,H,W,C] contiguous activations (fp32 or bf16).
"""
import ctypes

import torch

from . import _lib
from ._lib import F32, BF16, F16, ConvDesc, DeformHeadDesc, OffsetLevel, check, ptr, stream_handle

BN_EPS = 1e-5

# ---- optional per-call CUDA-event profiling (bench.py's roofline leg) -----------------------------
_PROF = None


def prof_begin():
    global _PROF
    _PROF = []


def prof_end():
    """-> list of (label, algorithmic_work, milliseconds); work is FLOPs for conv/deform, bytes for detect."""
    global _PROF
    rec, _PROF = _PROF, None
    torch.cuda.synchronize()
    return [(label, work, e0.elapsed_time(e1)) for (label, work, e0, e1) in rec]


def prof_take():
    """-> the raw records [(label, work, start event, end event)] and stops recording.  For records made while a CUDA graph
    was being captured (external events = event-record nodes of the graph): replay the graph, synchronize, then read
    ``e0.elapsed_time(e1)`` -- the kernels' own times inside the replay, without the host's launch overhead between two
    eager calls (which is what bounds the sub-20-microsecond launches of the small pyramid levels in eager mode)."""
    global _PROF
    rec, _PROF = _PROF, None
    return rec


class _Timed(object):
    def __init__(self, label, work):
        self.label, self.work = label, work

    def __enter__(self):
        if _PROF is not None:
            ext = torch.cuda.is_current_stream_capturing()
            self.e0 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e1 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _PROF is not None:
            self.e1.record()
            _PROF.append((self.label, self.work, self.e0, self.e1))
        return False


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError('unsupported dtype %s' % t.dtype)


def _torch_dt(code):
    return torch.float32 if code == F32 else torch.bfloat16


def _cuda(t, name):
    if not t.is_cuda:
        raise NotImplementedError('%s must be a CUDA tensor: tdrn_b200 has no CPU path' % name)
    return t.contiguous()


def conv_out(n, k, stride, pad, dil):
    return (n + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


class PackedConv(object):
    """Weights of one conv layer, BN-folded, packed for both kernels.

    w_f32  [kh*kw*Cin, Cout] fp32 (tap-major, then cin)      -> tdrn_conv2d (SIMT, fp32 accumulate)
    w_bf16 [Cout_pad, kh*kw*Cin_pad] bf16, K-major           -> tdrn_conv2d_tc (tcgen05); Cin_pad = Cin up to 64
    w_f16  the same packing in IEEE half (want_f16)           -> tdrn_conv2d_tc with in_dtype TDRN_F16
    w_x3   [Cout_pad, kh*kw*2*Cin] bf16 (W_hi | W_lo per tap)  -> tdrn_conv2d_tc with split3 (fp32-accurate tensor-core path)
    deconv (ConvTranspose2d k2 s2, weight [Cin,Cout,2,2]): w_f32 [Cin, 4*Cout] with n = (i*2+j)*Cout+co
    """

    def __init__(self, weight, bias=None, bn=None, stride=1, pad=0, dil=1, deconv=False, device='cuda',
                 want_bf16=True, want_x3=False, want_f16=False):
        w = weight.detach().double().cpu()
        b = bias.detach().double().cpu() if bias is not None else None
        self.deconv = deconv
        self.stride, self.pad, self.dil = stride, pad, dil
        if deconv:
            self.cin, self.cout, self.kh, self.kw = w.shape[0], w.shape[1], 2, 2
            assert tuple(w.shape[2:]) == (2, 2) and bn is None
            self.w_f32 = w.permute(0, 2, 3, 1).reshape(self.cin, 4 * self.cout).float().contiguous().to(device)
            # K-major rows n = (i*2+j)*Cout + co
            wk = w.permute(2, 3, 1, 0).reshape(4 * self.cout, self.cin)
        else:
            self.cout, self.cin, self.kh, self.kw = w.shape
            if bn is not None:                      # fold eval-mode BatchNorm2d (eps 1e-5) in float64
                gamma, beta, mean, var = [t.detach().double().cpu() for t in bn]
                scale = gamma / torch.sqrt(var + BN_EPS)
                w = w * scale.view(-1, 1, 1, 1)
                b = (b if b is not None else torch.zeros_like(mean)) - mean
                b = b * scale + beta
            self.w_f32 = w.permute(2, 3, 1, 0).reshape(self.kh * self.kw * self.cin, self.cout).float().contiguous().to(device)
            wk = w.permute(0, 2, 3, 1).reshape(self.cout, self.kh * self.kw * self.cin)
        self.bias = b.float().contiguous().to(device) if b is not None else None
        self.w_bf16 = None
        self.w_f16 = None
        self.w_x3 = None
        if want_x3 and self.cin % 64 == 0:
            # fp32-accurate tensor-core mode (tdrn_conv_desc.split3): [rows_pad][taps][W_hi | W_lo] bf16, where the fp32
            # weight (BN folded in float64, rounded to fp32 like the reference's parameters) is hi + lo to 16 mantissa bits
            rows = wk.shape[0]
            rows_pad = (rows + 15) // 16 * 16
            taps = wk.shape[1] // self.cin
            w32 = wk.float().reshape(rows, taps, self.cin)
            hi = w32.to(torch.bfloat16)
            lo = (w32 - hi.float()).to(torch.bfloat16)
            wp = torch.zeros(rows_pad, taps, 2 * self.cin, dtype=torch.bfloat16)
            wp[:rows] = torch.cat([hi, lo], 2)
            self.w_x3 = wp.reshape(rows_pad, taps * 2 * self.cin).contiguous().to(device)
        if want_bf16 and self.cin % 8 == 0 and self.cin >= 16:
            rows = wk.shape[0]
            rows_pad = (rows + 15) // 16 * 16
            cin_pad = (self.cin + 63) // 64 * 64                 # K padded per tap: the TMA channel box is 64 wide
            taps = wk.shape[1] // self.cin
            wp = torch.zeros(rows_pad, taps, cin_pad, dtype=torch.float64)
            wp[:rows, :, :self.cin] = wk.reshape(rows, taps, self.cin)
            self.w_bf16 = wp.reshape(rows_pad, taps * cin_pad).to(torch.bfloat16).contiguous().to(device)
            if want_f16:           # the same packing as IEEE half: operands of the TDRN_F16 convs (MobileNet trunk)
                self.w_f16 = wp.reshape(rows_pad, taps * cin_pad).clamp(-65504.0, 65504.0).to(torch.float16).contiguous().to(device)


def ensure_f16(pc):
    """Half-packed weights of ``pc`` (PackedConv.w_f16), made on first use from the fp32 copy when the layer was packed without them."""
    if pc.w_f16 is None and pc.w_bf16 is not None and not pc.deconv:
        rows_pad, kpad = pc.w_bf16.shape
        taps = pc.kh * pc.kw
        cin_pad = kpad // taps
        wp = torch.zeros(rows_pad, taps, cin_pad, dtype=torch.float32, device=pc.w_f32.device)
        wp[:pc.cout, :, :pc.cin] = pc.w_f32.t().reshape(pc.cout, taps, pc.cin)
        pc.w_f16 = wp.reshape(rows_pad, kpad).clamp(-65504.0, 65504.0).to(torch.float16).contiguous()
    return pc.w_f16


def conv_first_f16_ok(x_nchw, stride=2, cout=32):
    """Does the tensor-core stem (the only one with half output) tile this image?  Mirrors launch_conv_stem_tc (conv_stem_tc.cu)."""
    H, W = x_nchw.shape[2], x_nchw.shape[3]
    if not ((stride == 2 and cout == 32 and H % 2 == 0 and W % 2 == 0) or (stride == 1 and cout == 64)):
        return False
    Ho, Wo = conv_out(H, 3, stride, 1, 1), conv_out(W, 3, stride, 1, 1)
    tiles = (Wo % 64 == 0 and Ho % 2 == 0) or (Wo % 32 == 0 and Ho % 4 == 0) or (Wo % 16 == 0 and Ho % 8 == 0)
    return tiles and W % 4 == 0 and x_nchw.data_ptr() % 16 == 0


def split_bf16(x_nhwc):
    """fp32 NHWC [B,H,W,C] -> the split operand [B,H,W,2C] bf16 (hi | lo) of the fp32-accurate tensor-core convs."""
    x = _cuda(x_nhwc, 'input')
    assert x.dtype == torch.float32
    C = x.shape[-1]
    out = torch.empty(x.shape[:-1] + (2 * C,), dtype=torch.bfloat16, device=x.device)
    with _Timed('aux|split %d' % C, float(x.numel() * 8)):
        check(_lib.lib().tdrn_split_bf16(ptr(x), ptr(out), ctypes.c_longlong(x.numel() // C), C, stream_handle()), 'tdrn_split_bf16')
    return out


def conv2d(x_nhwc, pc, relu=False, out=None, out_dtype=None, residual=None, out_sb=None, out_sp=None,
           offsets=None, dg=0, use_tc=False, in_shape=None, in_sb=0, pool=False, label=None, work=None, split3=False,
           split_out=0):
    """out = act(conv(x) + bias (+ residual)).  ``out`` may be a view into a larger flat buffer, in which
    case out_sb/out_sp give the per-image and per-pixel strides (elements)."""
    if in_shape is not None:          # x is a strided view (e.g. one level of the flat [B,P,4] ARM output)
        x = x_nhwc
        B, H, W, Cin = in_shape
    else:
        x = _cuda(x_nhwc, 'input')
        B, H, W, Cin = x.shape
        if split3:
            Cin //= 2                                # x is the split tensor [B,H,W,2*Cin]
    assert Cin == pc.cin, (Cin, pc.cin)
    if pc.deconv:
        Ho, Wo = 2 * H, 2 * W
    else:
        Ho, Wo = conv_out(H, pc.kh, pc.stride, pc.pad, pc.dil), conv_out(W, pc.kw, pc.stride, pc.pad, pc.dil)
    Hs, Ws = (Ho // 2, Wo // 2) if pool else (Ho, Wo)        # stored map (after the fused 2x2 max-pool)
    if out is None and split_out:
        # the fp32 result leaves the kernel as the (hi | lo) bf16 operand of the next fp32-accurate conv: [B,Hs,Ws,2*Cout]
        out = torch.empty(B, Hs, Ws, 2 * pc.cout, dtype=torch.bfloat16, device=x.device)
        out_sb, out_sp = Hs * Ws * 2 * pc.cout, 2 * pc.cout
    if out is None:
        odt = out_dtype if out_dtype is not None else x.dtype
        out = torch.empty(B, Hs, Ws, pc.cout, dtype=odt, device=x.device)
    if out_sb is None:
        out_sb, out_sp = Hs * Ws * pc.cout, pc.cout
    d = ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=pc.cout, kh=pc.kh, kw=pc.kw, stride=pc.stride, pad=pc.pad,
                 dil=pc.dil, relu=int(relu), deconv2x2=int(pc.deconv), dg=dg, in_dtype=_dt(x),
                 out_dtype=_dt(out), out_sb=out_sb, out_sp=out_sp, in_sb=in_sb, pool2x2=int(pool), split3=int(split3), split_out=int(split_out))
    L = _lib.lib()
    flops = 2.0 * B * (H * W * 4 if pc.deconv else Ho * Wo * pc.kh * pc.kw) * Cin * pc.cout if work is None else work
    if use_tc:
        wt = pc.w_x3 if split3 else (ensure_f16(pc) if x.dtype == torch.float16 else pc.w_bf16)
        if wt is None or x.dtype not in (torch.bfloat16, torch.float16) or dg:
            raise _lib.TdrnError('tcgen05 conv needs bf16 (or half, with half-packed weights) input, Cin %% 8 == 0 and no offsets')
        with _Timed(label or '%s|%dx%d k%d d%d @%dx%d%s' % ('conv_tc_x3' if split3 else 'conv_tc', Cin, pc.cout, pc.kh, pc.dil, H, W,
                                                            ' deconv' if pc.deconv else ''), flops):
            check(L.tdrn_conv2d_tc(ctypes.byref(d), ptr(x), ptr(wt), ptr(pc.bias), ptr(residual), ptr(out),
                                   stream_handle()), 'tdrn_conv2d_tc')
    else:
        with _Timed('%s|%dx%d k%d s%d @%dx%d' % ('deform_simt' if dg else 'conv_simt', Cin, pc.cout, pc.kh, pc.stride, H, W), flops):
            check(L.tdrn_conv2d(ctypes.byref(d), ptr(x), ptr(pc.w_f32), ptr(pc.bias), ptr(residual), ptr(offsets),
                                ptr(out), stream_handle()), 'tdrn_conv2d')
    return out


def conv_first_split_ok(x_nchw, pc):
    """Can tdrn_conv_first write the (hi | lo) operand of the next fp32-accurate conv directly (TDRN_BF16_SPLIT)?"""
    H, W = x_nchw.shape[2], x_nchw.shape[3]
    tiles = (W % 64 == 0 and H % 2 == 0) or (W % 32 == 0 and H % 4 == 0) or (W % 16 == 0 and H % 8 == 0)
    return pc.cout == 64 and pc.stride == 1 and tiles and x_nchw.data_ptr() % 16 == 0


def conv_first(x_nchw, pc, relu, out_dtype, split=False):
    x = _cuda(x_nchw, 'input')
    if x.dtype != torch.float32:
        raise TypeError('network input must be float32 NCHW (reference boundary)')
    B, C, H, W = x.shape
    assert C == 3 and pc.kh == 3 and pc.pad == 1
    Ho, Wo = conv_out(H, 3, pc.stride, 1, 1), conv_out(W, 3, pc.stride, 1, 1)
    if split:
        out = torch.empty(B, Ho, Wo, 2 * pc.cout, dtype=torch.bfloat16, device=x.device)
        with _Timed('conv_first_x3', 2.0 * B * Ho * Wo * 27 * pc.cout):
            check(_lib.lib().tdrn_conv_first(ptr(x), ptr(pc.w_f32), ptr(pc.bias), ptr(out), B, H, W, pc.cout, pc.stride,
                                             int(relu), 2, stream_handle()), 'tdrn_conv_first')
        return out
    out = torch.empty(B, Ho, Wo, pc.cout, dtype=out_dtype, device=x.device)
    with _Timed('conv_first', 2.0 * B * Ho * Wo * 27 * pc.cout):
        check(_lib.lib().tdrn_conv_first(ptr(x), ptr(pc.w_f32), ptr(pc.bias), ptr(out), B, H, W, pc.cout, pc.stride,
                                         int(relu), _dt(out), stream_handle()), 'tdrn_conv_first')
    return out


def conv_stem_pair(x_nchw, pc1, pc2, relu=True, pool=True):
    """conv1_1 + conv1_2 (+ 2x2 max-pool) of the VGG trunk in one launch (tdrn_conv_stem_pair); returns None when the map
    does not tile (the caller then runs the two layers separately -- identical results)."""
    x = _cuda(x_nchw, 'input')
    if x.dtype != torch.float32:
        raise TypeError('network input must be float32 NCHW (reference boundary)')
    B, C, H, W = x.shape
    if (C != 3 or pc1.cout != 64 or pc2.cin != 64 or pc2.cout != 64 or pc2.w_bf16 is None or W % 8 or H % 16
            or pc1.stride != 1 or pc2.stride != 1 or pc2.kh != 3 or pc2.pad != 1 or pc2.dil != 1):
        return None
    out = torch.empty(B, H // 2 if pool else H, W // 2 if pool else W, 64, dtype=torch.bfloat16, device=x.device)
    flops = 2.0 * B * H * W * 64 * (27 + 9 * 64)
    with _Timed('conv_tc|3x64+64x64 k3 stem pair @%dx%d' % (H, W), flops):
        check(_lib.lib().tdrn_conv_stem_pair(ptr(x), ptr(pc1.w_f32), ptr(pc1.bias), ptr(pc2.w_bf16), ptr(pc2.bias), ptr(out),
                                             B, H, W, int(relu), int(relu), int(pool), stream_handle()), 'tdrn_conv_stem_pair')
    return out


class PackedDw(object):
    """Depthwise 3x3 weights [C,1,3,3] (+BN) -> [9, C] fp32."""

    def __init__(self, weight, bn, stride, device='cuda'):
        w = weight.detach().double().cpu()
        gamma, beta, mean, var = [t.detach().double().cpu() for t in bn]
        scale = gamma / torch.sqrt(var + BN_EPS)
        w = w * scale.view(-1, 1, 1, 1)
        self.c = w.shape[0]
        self.stride = stride
        self.w = w.reshape(self.c, 9).t().float().contiguous().to(device)
        self.bias = (beta - mean * scale).float().contiguous().to(device)


def dwconv3x3(x_nhwc, pd, relu=True, out_dtype=None):
    x = _cuda(x_nhwc, 'input')
    B, H, W, C = x.shape
    Ho, Wo = conv_out(H, 3, pd.stride, 1, 1), conv_out(W, 3, pd.stride, 1, 1)
    out = torch.empty(B, Ho, Wo, C, dtype=out_dtype or x.dtype, device=x.device)
    with _Timed('aux|dwconv3x3 %d s%d @%dx%d' % (C, pd.stride, H, W), float((x.numel() + out.numel()) * x.element_size())):
        check(_lib.lib().tdrn_dwconv3x3_io(ptr(x), ptr(pd.w), ptr(pd.bias), ptr(out), B, H, W, C, pd.stride, int(relu),
                                           _dt(x), _dt(out), stream_handle()), 'tdrn_dwconv3x3')
    return out


def conv_dwpw(x_nhwc, pd, pc, relu_dw=True, relu_pw=True):
    """One conv_dw block (depthwise 3x3 + pointwise 1x1, BN folded, ReLU after each) in one launch (tdrn_conv_dwpw):
    bit-identical to dwconv3x3 followed by conv2d(use_tc=True), without the HBM round trip of the depthwise output."""
    x = _cuda(x_nhwc, 'input')
    B, H, W, C = x.shape
    assert x.dtype == torch.bfloat16 and C == pd.c == pc.cin and pc.kh == 1 and pc.w_bf16 is not None
    Ho, Wo = conv_out(H, 3, pd.stride, 1, 1), conv_out(W, 3, pd.stride, 1, 1)
    out = torch.empty(B, Ho, Wo, pc.cout, dtype=torch.bfloat16, device=x.device)
    d = _lib.DwPwDesc(B=B, H=H, W=W, Cin=C, Cout=pc.cout, stride=pd.stride, relu_dw=int(relu_dw), relu_pw=int(relu_pw))
    with _Timed('conv_tc|dw3x3 s%d + %dx%d k1 @%dx%d' % (pd.stride, C, pc.cout, H, W), 2.0 * B * Ho * Wo * C * (pc.cout + 9)):
        check(_lib.lib().tdrn_conv_dwpw(ctypes.byref(d), ptr(x), ptr(pd.w), ptr(pd.bias), ptr(pc.w_bf16), ptr(pc.bias), ptr(out),
                                        stream_handle()), 'tdrn_conv_dwpw')
    return out


def maxpool2x2(x_nhwc, ceil_mode=False):
    x = _cuda(x_nhwc, 'input')
    B, H, W, C = x.shape
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if ceil_mode else (H // 2, W // 2)
    out = torch.empty(B, Ho, Wo, C, dtype=x.dtype, device=x.device)
    with _Timed('aux|maxpool %d @%dx%d' % (C, H, W), float((x.numel() + out.numel()) * x.element_size())):
        check(_lib.lib().tdrn_maxpool2x2(ptr(x), ptr(out), B, H, W, C, int(ceil_mode), _dt(x), stream_handle()),
              'tdrn_maxpool2x2')
    return out


def l2norm(x_nhwc, weight_f32, out_dtype=None):
    x = _cuda(x_nhwc, 'input')
    out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    C = x.shape[-1]
    with _Timed('aux|l2norm %d' % C, float(2 * x.numel() * x.element_size())):
        check(_lib.lib().tdrn_l2norm_io(ptr(x), ptr(weight_f32), ptr(out), ctypes.c_longlong(x.numel() // C), C, _dt(x), _dt(out),
                                        stream_handle()), 'tdrn_l2norm')
    return out


def l2norm_pool(x_nhwc, weight_f32):
    """-> (L2Norm(x), maxpool2x2(x)); one fused pass when the shape allows it (bf16, C in 256/512/1024, even H, W)."""
    x = _cuda(x_nhwc, 'input')
    B, H, W, C = x.shape
    if x.dtype != torch.bfloat16 or C not in (256, 512, 1024) or H % 2 or W % 2:
        return l2norm(x, weight_f32), maxpool2x2(x, False)
    out_n = torch.empty_like(x)
    out_p = torch.empty(B, H // 2, W // 2, C, dtype=x.dtype, device=x.device)
    with _Timed('aux|l2norm+pool %d @%dx%d' % (C, H, W), float((2 * x.numel() + out_p.numel()) * 2)):
        check(_lib.lib().tdrn_l2norm_pool2x2(ptr(x), ptr(weight_f32), ptr(out_n), ptr(out_p), B, H, W, C, _dt(x),
                                             stream_handle()), 'tdrn_l2norm_pool2x2')
    return out_n, out_p


def softmax_rows(x, out=None):
    x = _cuda(x, 'input')
    assert x.dtype == torch.float32 and x.dim() == 2
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().tdrn_softmax(ptr(x), ptr(out), ctypes.c_longlong(x.shape[0]), x.shape[1], stream_handle()),
          'tdrn_softmax')
    return out


def nhwc_to_nchw_f32(x_nhwc):
    x = _cuda(x_nhwc, 'input')
    B, H, W, C = x.shape
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=x.device)
    with _Timed('aux|to_nchw %d @%dx%d' % (C, H, W), float(x.numel() * x.element_size() + out.numel() * 4)):
        check(_lib.lib().tdrn_nhwc_to_nchw_f32(ptr(x), ptr(out), B, H, W, C, _dt(x), stream_handle()),
              'tdrn_nhwc_to_nchw_f32')
    return out


def nchw_f32_to_nhwc(x_nchw, dtype):
    x = _cuda(x_nchw, 'input')
    assert x.dtype == torch.float32
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, C, dtype=dtype, device=x.device)
    check(_lib.lib().tdrn_nchw_f32_to_nhwc(ptr(x), ptr(out), B, C, H, W, _dt(out), stream_handle()),
          'tdrn_nchw_f32_to_nhwc')
    return out


def offset_convs(arm_loc, sizes, lv_off, w1, b1, w2=None, b2=None, want_nchw=True):
    """All levels' `offset.k` (and `offset2.k`) 1x1 convs on the flattened ARM regression [B,P,4] in one launch.
    sizes [(H, W)], lv_off [prior offset], w1/w2 lists of [c,12] fp32 CUDA tensors, b1/b2 lists of [c] or None.
    -> (offsets NHWC fp32 list, offsets2 NHWC list (or []), offsets NCHW list (or None))"""
    arm = _cuda(arm_loc, 'arm_loc')
    assert arm.dtype == torch.float32
    B, P, _ = arm.shape
    n = len(sizes)
    c1 = w1[0].shape[0]
    c2 = w2[0].shape[0] if w2 else 0
    o1 = [torch.empty(B, h, w, c1, dtype=torch.float32, device=arm.device) for h, w in sizes]
    o2 = [torch.empty(B, h, w, c2, dtype=torch.float32, device=arm.device) for h, w in sizes] if c2 else []
    on = [torch.empty(B, c1, h, w, dtype=torch.float32, device=arm.device) for h, w in sizes] if want_nchw else None
    pv = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    lv = (OffsetLevel * n)()
    for k, (h, w) in enumerate(sizes):
        lv[k] = OffsetLevel(H=h, W=w, prior_off=int(lv_off[k]), w1=pv(w1[k]), b1=pv(b1[k] if b1 else None),
                            w2=pv(w2[k] if c2 else None), b2=pv(b2[k] if (c2 and b2) else None),
                            out1=pv(o1[k]), out2=pv(o2[k] if c2 else None), out1_nchw=pv(on[k] if want_nchw else None))
    work = 2.0 * B * sum(h * w for h, w in sizes) * 12 * (c1 + c2)
    with _Timed('conv_simt|offset convs, %d levels 12x%d+%d' % (n, c1, c2), work):
        check(_lib.lib().tdrn_offset_convs(ptr(arm), B, P, n, lv, c1, c2, stream_handle()), 'tdrn_offset_convs')
    return o1, o2, on


def deform_conv_nchw(input, offset, weight, stride, pad, dil, dg):
    """The reference operator boundary: deform_conv_forward_cuda (NCHW fp32)."""
    for name, t in (('input', input), ('offset', offset), ('weight', weight)):
        if not t.is_cuda or t.dtype != torch.float32:
            raise NotImplementedError('%s must be a torch.cuda.FloatTensor' % name)   # networks.py:632-640
    input, offset, weight = input.contiguous(), offset.contiguous(), weight.contiguous()
    if input.dim() != 4:
        raise ValueError('Expected 4D tensor as input, got {}D tensor instead.'.format(input.dim()))
    B, C, H, W = input.shape
    Cout, Cin, kh, kw = weight.shape
    Ho, Wo = conv_out(H, kh, stride[0], pad[0], dil[0]), conv_out(W, kw, stride[1], pad[1], dil[1])
    if Ho <= 0 or Wo <= 0:
        raise ValueError('convolution input is too small (output would be {}x{}x{}x{})'.format(B, Cout, Ho, Wo))
    if Cin != C:
        raise RuntimeError('invalid number of input planes, expected: %d, but got: %d' % (Cin, C))
    if tuple(offset.shape) != (B, dg * 2 * kh * kw, Ho, Wo):
        raise RuntimeError('invalid offset shape %s, expected %s' % (tuple(offset.shape), (B, dg * 2 * kh * kw, Ho, Wo)))
    out = torch.empty(B, Cout, Ho, Wo, dtype=torch.float32, device=input.device)
    check(_lib.lib().tdrn_deform_conv_forward(ptr(input), ptr(weight), ptr(offset), ptr(out), B, C, H, W, Cout,
                                              kw, kh, stride[1], stride[0], pad[1], pad[0], dil[0], dil[1], dg,
                                              stream_handle()), 'tdrn_deform_conv_forward')
    return out


def pack_deform_head_weight(wcat, device='cuda'):
    """[N, Cin, kh, kw] (loc rows then conf rows) -> bf16 [N_pad16, K] K-major for tdrn_deform_head, with
    k = (cb * taps + tap) * 64 + c: channel-block-major so all taps of one 64-channel slab are consecutive."""
    w = wcat.detach().double().cpu()
    n, cin, kh, kw = w.shape
    assert cin % 64 == 0
    wk = w.permute(0, 2, 3, 1).reshape(n, kh * kw, cin // 64, 64).permute(0, 2, 1, 3).reshape(n, kh * kw * cin)
    n_pad = (n + 15) // 16 * 16
    wp = torch.zeros(n_pad, wk.shape[1], dtype=torch.float64)
    wp[:n] = wk
    return wp.to(torch.bfloat16).contiguous().to(device)


def deform_head(feat_nhwc, offsets, w_bf16, num_classes, dg, kh, pad, loc_out, conf_out, P, prior_off,
                offsets2=None, w2_bf16=None, kh2=0, pad2=0, softmax=True):
    x = _cuda(feat_nhwc, 'feat')
    B, H, W, Cin = x.shape
    d = DeformHeadDesc(B=B, H=H, W=W, Cin=Cin, num_classes=num_classes, dg=dg, kh=kh, pad=pad, kh2=kh2, pad2=pad2,
                       P=P, prior_off=prior_off, softmax=int(softmax))
    flops = 2.0 * B * H * W * (12 + 3 * num_classes) * Cin * (kh * kh + kh2 * kh2)
    with _Timed('deform_head_tc|%d @%dx%d k%d+%d' % (Cin, H, W, kh, kh2), flops):
        check(_lib.lib().tdrn_deform_head(ctypes.byref(d), ptr(x), ptr(offsets), ptr(w_bf16), ptr(offsets2),
                                          ptr(w2_bf16), ptr(loc_out), ptr(conf_out), stream_handle()), 'tdrn_deform_head')


def pack_deform_proj_weight(wcat, wcat2=None, device='cuda', x3=False):
    """Per-tap projection weights of the "project, then sample" head (tdrn_deform_head_sample).

    wcat [N, Cin, kh, kw] (loc rows then conf rows; wcat2: the optional second, 5x5, head) -> a 1x1 PackedConv with
    Cout = taps * n_pad rows, row t * n_pad + o = W[o, :, tap t] (taps of head 1 first), n_pad = N up to a multiple of 8."""
    ws = [wcat.detach().double().cpu()] + ([wcat2.detach().double().cpu()] if wcat2 is not None else [])
    n, cin = ws[0].shape[:2]
    n_pad = (n + 15) // 16 * 16 if x3 else (n + 7) // 8 * 8     # x3 (fp32-accurate heads): groups of g = n_pad (hi | lo) channels
    rows = []
    for w in ws:
        kh, kw = w.shape[2:]
        blk = torch.zeros(kh * kw, n_pad, cin, dtype=torch.float64)
        blk[:, :n] = w.permute(2, 3, 0, 1).reshape(kh * kw, n, cin)
        rows.append(blk.reshape(kh * kw * n_pad, cin))
    wp = torch.cat(rows, 0)
    return PackedConv(wp.view(wp.shape[0], cin, 1, 1), device=device, want_bf16=not x3, want_x3=x3), n_pad


def _deform_chunk_bytes():
    import os
    return int(float(os.environ.get('TDRN_DEFORM_CHUNK_MB', '1024')) * (1 << 20))


def half_projections():
    """OPT-IN (TDRN_PROJ_F16=1): store the per-tap projections of the 16-bit deformable heads as IEEE half instead of bf16 (same
    bytes; tdrn_deform_head_desc.split = 2).  Measured on B200 with the oracle's offsets given (scripts/mobile_half_check.py,
    profiles/r03i_half_check.txt): no gain -- VGG conf 8.2e-3 (bf16) / 8.8e-3 (half), MobileNet 1.60e-2 / 1.69e-2: the rounding of
    the projections averages out over the 34 x 4 sampled values of a row, the error of the heads comes from their bf16 inputs."""
    import os
    return os.environ.get('TDRN_PROJ_F16', '0') == '1'


def deform_head_projected(feat_nhwc, offsets, pc_proj, n_pad, num_classes, kh, pad, loc_out, conf_out, P, prior_off,
                          offsets2=None, kh2=0, pad2=0, softmax=True, split=False):
    """Same contract as deform_head (dg = 1) in two launches per image chunk: dense per-tap projection on tcgen05
    (1x1 conv, bf16 out), then the bilinear sampler over the projections.  The batch is split into image chunks only
    when the projection buffer would exceed TDRN_DEFORM_CHUNK_MB (default 1024 MB; 278 MB at b32 / 40x40 / VOC-21).
    Measured on B200: L2-sized chunks (48 MB) lose more to per-launch tails than they gain in L2 residency."""
    x = _cuda(feat_nhwc, 'feat')
    B, H, W, Cin = x.shape
    taps = kh * kh + kh2 * kh2
    assert pc_proj.cout == taps * n_pad and pc_proj.cin == Cin
    # split (fp32-accurate heads): x fp32 -> (hi | lo) operand, x3 GEMM, projections stored as (hi | lo) pairs per tap
    width = 2 * n_pad if split else n_pad
    per_img = H * W * taps * width * 2
    nb = max(1, min(B, _deform_chunk_bytes() // per_img))
    half = not split and half_projections()
    y = torch.empty(nb, H, W, taps * width, dtype=torch.float16 if half else torch.bfloat16, device=x.device)
    flops = 2.0 * H * W * (12 + 3 * num_classes) * Cin * taps
    L = _lib.lib()
    xin = split_bf16(x) if split else x
    tag = 'deform_head_x3' if split else 'deform_head_tc'
    for b0 in range(0, B, nb):
        n = min(nb, B - b0)
        if split:
            conv2d(xin[b0:b0 + n], pc_proj, use_tc=True, out=y[:n], split3=True, split_out=n_pad, out_sb=H * W * taps * width,
                   out_sp=taps * width, label='%s|%d @%dx%d k%d+%d project' % (tag, Cin, H, W, kh, kh2), work=flops * n)
        else:
            conv2d(xin[b0:b0 + n], pc_proj, use_tc=True, out=y[:n],
                   label='%s|%d @%dx%d k%d+%d project' % (tag, Cin, H, W, kh, kh2), work=flops * n)
        d = DeformHeadDesc(B=n, H=H, W=W, Cin=Cin, num_classes=num_classes, dg=1, kh=kh, pad=pad, kh2=kh2, pad2=pad2,
                           P=P, prior_off=prior_off, softmax=int(softmax), split=2 if half else int(split))
        with _Timed('%s|%d @%dx%d k%d+%d sample' % (tag, Cin, H, W, kh, kh2), 0.0):
            check(L.tdrn_deform_head_sample(ctypes.byref(d), ptr(y), width, ptr(offsets[b0:b0 + n]),
                                            ptr(offsets2[b0:b0 + n]) if offsets2 is not None else None,
                                            ptr(loc_out[b0:b0 + n]), ptr(conf_out[b0:b0 + n]), stream_handle()),
                  'tdrn_deform_head_sample')


def deform_project(feat_nhwc, pc_proj, n_pad, kh, kh2, num_classes, split=False, xs=None):
    """First half of deform_head_projected for a whole batch: the per-tap projections [B,H,W,taps*width] bf16.
    ``xs``: the (hi | lo) operand of ``feat_nhwc`` when the caller already has it (split mode)."""
    x = _cuda(feat_nhwc, 'feat')
    B, H, W, Cin = x.shape
    taps = kh * kh + kh2 * kh2
    width = 2 * n_pad if split else n_pad
    y = torch.empty(B, H, W, taps * width, dtype=torch.float16 if (not split and half_projections()) else torch.bfloat16, device=x.device)
    flops = 2.0 * B * H * W * (12 + 3 * num_classes) * Cin * taps
    tag = 'deform_head_x3' if split else 'deform_head_tc'
    label = '%s|%d @%dx%d k%d+%d project' % (tag, Cin, H, W, kh, kh2)
    if split:
        conv2d(xs if xs is not None else split_bf16(x), pc_proj, use_tc=True, out=y, split3=True, split_out=n_pad,
               out_sb=H * W * taps * width, out_sp=taps * width, label=label, work=flops)
    else:
        conv2d(x, pc_proj, use_tc=True, out=y, label=label, work=flops)
    return y


def deform_sample_group(projs, shapes, n_pad, num_classes, kh, pad, offsets, loc_out, conf_out, P, prior_offs,
                        offsets2=None, kh2=0, pad2=0, softmax=True, split=False):
    """Second half for ALL pyramid levels in one launch (tdrn_deform_head_sample_group).  projs[k] from deform_project,
    shapes[k] = (B, H, W, Cin) of level k's feature map."""
    n = len(projs)
    width = 2 * n_pad if split else n_pad
    descs = (DeformHeadDesc * n)()
    for k, (B, H, W, Cin) in enumerate(shapes):
        descs[k] = DeformHeadDesc(B=B, H=H, W=W, Cin=Cin, num_classes=num_classes, dg=1, kh=kh, pad=pad, kh2=kh2, pad2=pad2,
                                  P=P, prior_off=int(prior_offs[k]), softmax=int(softmax),
                                  split=2 if projs[k].dtype == torch.float16 else int(split))
    vp = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
    with _Timed('%s|all %d levels k%d+%d sample' % ('deform_head_x3' if split else 'deform_head_tc', n, kh, kh2), 0.0):
        check(_lib.lib().tdrn_deform_head_sample_group(n, descs, vp(projs), width, vp(offsets),
                                                       vp(offsets2) if offsets2 else None, ptr(loc_out), ptr(conf_out),
                                                       stream_handle()), 'tdrn_deform_head_sample_group')


def preprocess(frames_u8, size, mean, swap_rb=False, out=None, flip_lr=False):
    """[B,Hs,Ws,3] uint8 CUDA frames (cv2 channel order) -> [B,3,size,size] fp32 NCHW network input:
    base_transform (data/__init__.py:7-12) + optional channel swap + HWC->CHW, one launch (tdrn_preprocess)."""
    f = _cuda(frames_u8, 'frames')
    if f.dtype != torch.uint8 or f.dim() != 4 or f.shape[3] != 3:
        raise TypeError('frames must be a uint8 tensor [B,H,W,3]')
    B, Hs, Ws, _ = f.shape
    if out is None:
        out = torch.empty(B, 3, size, size, dtype=torch.float32, device=f.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    with _Timed('preprocess|%dx%d->%d' % (Hs, Ws, size), float(B * (Hs * Ws * 3 + 3 * size * size * 4))):
        check(_lib.lib().tdrn_preprocess(ptr(f), B, Hs, Ws, int(size), m, int(bool(swap_rb)), int(bool(flip_lr)), ptr(out),
                                         stream_handle()),
              'tdrn_preprocess')
    return out


def multiscale_vote(dets, flips, rules, rule_thrs, w, h, vote_thresh=0.45):
    """dets [K,C,top_k,5] CUDA (Detect outputs of the K passes of one image) -> (rows [C,K*top_k,5], counts [C]) on the
    device: per-class gather + bbox_vote of multi_eval.py:453-494,557-640 (tdrn_multiscale_vote)."""
    d = _cuda(dets, 'dets').float()
    K, C, top_k, five = d.shape
    assert five == 5 and len(flips) == K and len(rules) == K and len(rule_thrs) == K
    dev = d.device
    fl = torch.tensor([int(v) for v in flips], dtype=torch.int32, device=dev)
    ru = torch.tensor([int(v) for v in rules], dtype=torch.int32, device=dev)
    rt = torch.tensor([float(v) for v in rule_thrs], dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = torch.empty(L.tdrn_multiscale_vote_workspace_bytes(K, C, top_k), dtype=torch.uint8, device=dev)
    out = torch.zeros(C, K * top_k, 5, dtype=torch.float32, device=dev)
    cnt = torch.zeros(C, dtype=torch.int32, device=dev)
    import numpy as np
    check(L.tdrn_multiscale_vote(ptr(d), ptr(fl), ptr(ru), ptr(rt), K, C, top_k, ctypes.c_float(float(w)), ctypes.c_float(float(h)),
                                 ctypes.c_float(float(np.float32(vote_thresh))), ptr(out), ptr(cnt), K * top_k, ptr(ws),
                                 ctypes.c_size_t(ws.numel()), stream_handle()), 'tdrn_multiscale_vote')
    return out, cnt


def decode(loc, priors, arm_loc=None):
    loc, priors = _cuda(loc, 'loc').float(), _cuda(priors, 'priors').float()
    B, P, _ = loc.shape
    arm = _cuda(arm_loc, 'arm_loc').float() if arm_loc is not None else None
    out = torch.empty(B, P, 4, dtype=torch.float32, device=loc.device)
    check(_lib.lib().tdrn_decode(ptr(loc), ptr(priors), ptr(arm), B, P, ptr(out), stream_handle()), 'tdrn_decode')
    return out


_ws_cache = {}
_ws_retired = []          # outgrown workspaces stay allocated: a captured CUDA graph may hold their raw pointers


def _workspace(nbytes, device):
    """Scratch memory for detect / nms_device, one block per (device, stream): calls in flight on different streams
    (two graph instances replaying concurrently) must never share scratch memory.

    Lifetime rules (a captured CUDA graph bakes the raw pointer in):
      * a block that is outgrown is RETIRED, never freed, so graphs captured earlier on that stream keep writing into
        memory that is still theirs;
      * a block is never allocated or grown while the stream is capturing (it would live in that graph's private pool
        and be handed to eager calls and other captures afterwards): warm the call up eagerly first -- this raises;
      * a graph captured before a growth keeps using the retired block; re-capture it if it should use the new one."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.TdrnError('detect/nms workspace of %d bytes would be allocated during CUDA-graph capture; run the '
                                 'same call once eagerly on this stream before capturing' % nbytes)
        if ws is not None:
            _ws_retired.append(ws)
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def detect(loc, conf, priors, arm_loc, scale, num_classes, top_k, conf_thresh, nms_thresh, out=None):
    loc = _cuda(loc, 'loc_data').float()
    conf = _cuda(conf, 'conf_data').float()
    priors = _cuda(priors, 'prior_data').float()
    arm = _cuda(arm_loc, 'arm_loc_data').float() if arm_loc is not None else None
    B, P = loc.shape[0], priors.shape[0]
    L = _lib.lib()
    nbytes = L.tdrn_detect_workspace_bytes(B, P, num_classes, top_k)
    ws = _workspace(nbytes, loc.device)
    if out is None:
        out = torch.empty(B, num_classes, top_k, 5, dtype=torch.float32, device=loc.device)
    sc = (ctypes.c_float * 4)(*[float(v) for v in scale])
    # algorithmic bytes: arm_loc + odm_loc + conf in, [C,top_k,5] out per frame, priors once
    nbytes_alg = B * (P * 16 * (2 if arm is not None else 1) + P * num_classes * 4 + num_classes * top_k * 20) + P * 16
    with _Timed('detect', float(nbytes_alg)):
        check(L.tdrn_detect(ptr(loc), ptr(conf), ptr(priors), ptr(arm), sc, B, P, num_classes, top_k,
                            ctypes.c_float(conf_thresh), ctypes.c_double(nms_thresh), ptr(out), ptr(ws),
                            ctypes.c_size_t(ws.numel()), stream_handle()), 'tdrn_detect')
    return out


def nms_device(dets, thresh, max_keep=0):
    """dets [n,5] CUDA fp32 -> (keep int32 [n], num_keep int32 [1]) on device."""
    dets = _cuda(dets, 'dets').float()
    n = dets.shape[0]
    L = _lib.lib()
    ws = _workspace(L.tdrn_nms_workspace_bytes(n), dets.device)
    keep = torch.empty(max(n, 1), dtype=torch.int32, device=dets.device)
    num = torch.zeros(1, dtype=torch.int32, device=dets.device)
    check(L.tdrn_nms(ptr(dets), n, ctypes.c_double(thresh), max_keep, ptr(keep), ptr(num), ptr(ws),
                     ctypes.c_size_t(ws.numel()), stream_handle()), 'tdrn_nms')
    return keep, num


def prior_box(cfg):
    """PriorBox.forward via the C ABI (host) -> CPU fp32 tensor [P,4]."""
    n = len(cfg['feature_maps'])
    ia = lambda v: (ctypes.c_int * len(v))(*[int(t) for t in v])
    da = lambda v: (ctypes.c_double * len(v))(*[float(t) for t in v])       # sizes / ratios may be fractional (SSD-512, MOT_300)
    for name in ('feature_maps', 'steps'):
        if any(int(t) != t for t in cfg[name]):
            raise ValueError('%s must be integers' % name)
    ars, n_ar = [], []
    for a in cfg['aspect_ratios']:
        if len(a) > 4:
            raise ValueError('at most 4 aspect ratios per level are supported')
        ars += [float(t) for t in a] + [0.0] * (4 - len(a))
        n_ar.append(len(a))
    mx = da(cfg['max_sizes']) if len(cfg['max_sizes']) else None
    num = ctypes.c_int(0)
    L = _lib.lib()
    args = (int(cfg['min_dim']), n, ia(cfg['feature_maps']), ia(cfg['steps']), da(cfg['min_sizes']), mx, ia(n_ar),
            da(ars), int(bool(cfg['flip'])), int(bool(cfg['clip'])))
    check(L.tdrn_prior_box(*args, None, ctypes.byref(num)), 'tdrn_prior_box')
    out = torch.empty(num.value, 4, dtype=torch.float32)
    check(L.tdrn_prior_box(*args, ptr(out), ctypes.byref(num)), 'tdrn_prior_box')
    return out
