"""Pure-PyTorch (CPU, fp32) restatement of the reference deformable convolution forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line:
  * sampler  ``deformable_im2col_bilinear``   utils/deformconv/deform_conv_cuda_kernel.cu:16-51
  * im2col   ``deformable_im2col_gpu_kernel`` utils/deformconv/deform_conv_cuda_kernel.cu:157-208
  * host     ``deform_conv_forward_cuda``     utils/deformconv/deform_conv_cuda.c:141-193
    (zero output, columns[C*kh*kw, Ho*Wo], SGEMM with weight[Cout, C*kh*kw], no bias)
  * Python   ``ConvOffset2dFunction._output_size`` model/networks.py:684-697

Semantics that differ from torchvision.ops.deform_conv2d:
  * a sample is used only when 0 <= h < H and 0 <= w < W   (.cu:195) -- points in (-1, 0) give 0;
  * when floor(h) >= H-1 the row is clamped to H-1 with lh = 0 (.cu:25-30), same for w (.cu:32-37),
    i.e. points in [H-1, H) replicate the last row/column.

Parity status for this function: pinned on the GPU box against the reference's OWN CUDA kernel file, compiled
unmodified by oracle/build_ref.py and run in tests/test_gpu_ref_native.py (im2col columns within 4 ulp -- the
reference is built with -fmad=true -- and identical zero patterns for the border rules); on CPU it is cross-checked
against F.conv2d (zero offsets), torchvision (interior points) and the scalar C restatement in oracle/c/oracle.c.
"""
import torch


def output_size(h, w, kh, kw, stride, pad, dil):
    """model/networks.py:684-697 / deform_conv_cuda.c:131-134."""
    ho = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    wo = (w + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    if ho <= 0 or wo <= 0:
        raise ValueError("convolution input is too small (output would be {}x{})".format(ho, wo))
    return ho, wo


def deform_im2col(inp, offset, kh, kw, stride, pad, dil, dg):
    """columns[C*kh*kw, Ho*Wo] for ONE sample.  inp [C,H,W] fp32, offset [dg*2*kh*kw, Ho, Wo] fp32."""
    c, h, w = inp.shape
    ho, wo = output_size(h, w, kh, kw, stride, pad, dil)
    assert offset.shape == (dg * 2 * kh * kw, ho, wo), (offset.shape, (dg * 2 * kh * kw, ho, wo))
    assert c % dg == 0
    cpg = c // dg
    f32 = torch.float32
    inp = inp.to(f32)
    offset = offset.to(f32)

    ys = torch.arange(ho, dtype=torch.int64).view(1, ho, 1)
    xs = torch.arange(wo, dtype=torch.int64).view(1, 1, wo)
    h_in = ys * stride - pad                      # .cu:177
    w_in = xs * stride - pad                      # .cu:178
    ti = (torch.arange(kh * kw, dtype=torch.int64) // kw).view(-1, 1, 1)   # tap row i
    tj = (torch.arange(kh * kw, dtype=torch.int64) % kw).view(-1, 1, 1)    # tap col j

    cols = torch.empty(c, kh * kw, ho * wo, dtype=f32)
    flat = inp.reshape(c, h * w)
    for g in range(dg):
        off = offset[g * 2 * kh * kw:(g + 1) * 2 * kh * kw].view(kh * kw, 2, ho, wo)
        off_h = off[:, 0]                          # channel 2*(i*kw+j)     .cu:187-188,192
        off_w = off[:, 1]                          # channel 2*(i*kw+j)+1   .cu:189-191,193
        # .cu:195-196: integer part is summed in int, then added to the float offset
        h_im = (h_in + ti * dil).to(f32) + off_h
        w_im = (w_in + tj * dil).to(f32) + off_w
        valid = (h_im >= 0) & (w_im >= 0) & (h_im < h) & (w_im < w)      # .cu:197
        map_h = (ti * dil).to(f32) + off_h         # .cu:198
        map_w = (tj * dil).to(f32) + off_w         # .cu:199
        cur_h = (h - h_in).expand_as(map_h)        # .cu:200
        cur_w = (w - w_in).expand_as(map_w)        # .cu:201
        # sampler .cu:21-37 (coordinates relative to (h_in, w_in))
        h_low = torch.floor(map_h).to(torch.int64)
        w_low = torch.floor(map_w).to(torch.int64)
        clamp_h = h_low >= cur_h - 1
        clamp_w = w_low >= cur_w - 1
        h_low = torch.where(clamp_h, cur_h - 1, h_low)
        w_low = torch.where(clamp_w, cur_w - 1, w_low)
        h_high = torch.where(clamp_h, h_low, h_low + 1)
        w_high = torch.where(clamp_w, w_low, w_low + 1)
        hq = torch.where(clamp_h, h_low.to(f32), map_h)
        wq = torch.where(clamp_w, w_low.to(f32), map_w)
        lh = hq - h_low.to(f32)                    # .cu:39-41
        lw = wq - w_low.to(f32)
        hh = 1 - lh
        hw = 1 - lw
        w1, w2, w3, w4 = hh * hw, hh * lw, lh * hw, lh * lw                # .cu:47

        def corner(hr, wr):
            ha = (h_in + hr).clamp(0, h - 1)       # absolute row; clamp only guards invalid samples
            wa = (w_in + wr).clamp(0, w - 1)
            idx = (ha * w + wa).reshape(-1)
            return flat[g * cpg:(g + 1) * cpg].index_select(1, idx).view(cpg, kh * kw, ho * wo)

        v1 = corner(h_low, w_low)                  # .cu:43-46
        v2 = corner(h_low, w_high)
        v3 = corner(h_high, w_low)
        v4 = corner(h_high, w_high)
        r = lambda t: t.reshape(1, kh * kw, ho * wo)
        val = r(w1) * v1 + r(w2) * v2 + r(w3) * v3 + r(w4) * v4            # .cu:49
        val = torch.where(r(valid), val, torch.zeros((), dtype=f32))
        cols[g * cpg:(g + 1) * cpg] = val
    return cols.view(c * kh * kw, ho * wo)         # row index c*kh*kw + i*kw + j   .cu:172,180-181,204


def deform_conv_forward(inp, offset, weight, stride=1, pad=0, dil=1, dg=1):
    """out[B,Cout,Ho,Wo] = weight[Cout, C*kh*kw] @ columns   (deform_conv_cuda.c:157-193; no bias)."""
    assert inp.dim() == 4, "Expected 4D tensor as input"           # model/networks.py:608-611
    b, c, h, w = inp.shape
    cout, cin, kh, kw = weight.shape
    assert cin == c
    ho, wo = output_size(h, w, kh, kw, stride, pad, dil)
    assert offset.shape[0] == b, "invalid batch size of offset"    # deform_conv_cuda.c:136
    out = torch.empty(b, cout, ho, wo, dtype=torch.float32)
    wmat = weight.to(torch.float32).reshape(cout, c * kh * kw)
    for n in range(b):                                              # deform_conv_cuda.c:157
        cols = deform_im2col(inp[n], offset[n], kh, kw, stride, pad, dil, dg)
        out[n] = (wmat @ cols).view(cout, ho, wo)
    return out


def conv_offset2d(input, offset, weight, stride=1, padding=0, dilation=1, deform_groups=1):
    """Drop-in for model.networks.conv_offset2d (model/networks.py:600-615) used by ref_shim."""
    from torch.nn.modules.utils import _pair
    s, p, d = _pair(stride), _pair(padding), _pair(dilation)
    assert s[0] == s[1] and p[0] == p[1] and d[0] == d[1], "oracle handles square stride/pad/dilation"
    if input is not None and input.dim() != 4:
        raise ValueError("Expected 4D tensor as input, got {}D tensor instead.".format(input.dim()))
    return deform_conv_forward(input, offset, weight, s[0], p[0], d[0], deform_groups)
