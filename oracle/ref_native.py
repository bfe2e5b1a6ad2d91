"""ctypes loader of oracle/_ref/libtdrn_ref_native.so: the reference's OWN CUDA kernels (deformable im2col,
utils/deformconv/deform_conv_cuda_kernel.cu; GPU NMS, utils/nms/nms_kernel.cu) compiled unmodified by
oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY -- needs a GPU; used by tests/test_gpu_ref_native.py to pin the
restatements in oracle/ (and through them the product kernels) against the real reference.
"""
import ctypes
import os

import numpy as np

from . import build_ref

_lib = None


def available():
    return os.path.exists(build_ref.OUT)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError('%s not built (python -m oracle.build_ref in the build container)' % build_ref.OUT)
        _lib = ctypes.CDLL(build_ref.OUT)
    return _lib


def deformable_im2col(inp, offset, kh, kw, stride, pad, dil, dg):
    """One sample through the reference launcher (deform_conv_cuda_kernel.cu:211-238).
    inp [C,H,W], offset [dg*2*kh*kw,Ho,Wo] CUDA fp32 -> columns [C*kh*kw, Ho*Wo] CUDA fp32."""
    import torch
    assert inp.is_cuda and offset.is_cuda and inp.dtype == torch.float32 and offset.dtype == torch.float32
    inp, offset = inp.contiguous(), offset.contiguous()
    c, h, w = inp.shape
    ho = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    wo = (w + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    assert tuple(offset.shape) == (dg * 2 * kh * kw, ho, wo)
    col = torch.empty(c * kh * kw, ho * wo, dtype=torch.float32, device=inp.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib().ref_deformable_im2col(ctypes.c_void_p(inp.data_ptr()), ctypes.c_void_p(offset.data_ptr()), c, h, w, kh, kw,
                                     pad, pad, stride, stride, dil, dil, dg, ctypes.c_void_p(col.data_ptr()), st)
    if rc != 0:
        raise RuntimeError('reference deformable_im2col: cuda error %d' % rc)
    return col


def deform_conv_forward(inp, offset, weight, stride=1, pad=0, dil=1, dg=1):
    """The reference's host loop (deform_conv_cuda.c:157-193) around its own im2col kernel; the SGEMM
    (THCudaBlas_Sgemm, :185-192) is torch.matmul in fp32 (cuBLAS, TF32 disabled), output zero-initialised, no bias."""
    import torch
    b, c, h, w = inp.shape
    cout, _, kh, kw = weight.shape
    wmat = weight.reshape(cout, c * kh * kw).float()
    outs = []
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for n in range(b):
            col = deformable_im2col(inp[n], offset[n], kh, kw, stride, pad, dil, dg)
            outs.append(wmat @ col)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    ho = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    return torch.stack(outs, 0).view(b, cout, ho, -1)


def gpu_nms(dets, thresh, device_id=0):
    """utils/nms/gpu_nms.pyx:16-31: sort by score descending on the host, call `_nms` on the sorted boxes,
    map the kept positions back through the order.  dets [N,5] float32 ndarray -> list of indices."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return []
    order = dets[:, 4].argsort()[::-1]
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(n, dtype=np.int32)
    num = ctypes.c_int(0)
    rc = lib().ref_gpu_nms(keep.ctypes.data_as(ctypes.c_void_p), ctypes.byref(num),
                           sorted_dets.ctypes.data_as(ctypes.c_void_p), n, 5, ctypes.c_float(thresh), device_id)
    if rc != 0:
        raise RuntimeError('reference _nms: cuda error %d' % rc)
    return list(order[keep[:num.value]])
