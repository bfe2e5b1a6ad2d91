"""ctypes loader for libtdrn_b200.so (the C ABI declared in include/tdrn_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Device memory, streams and torch.distributed come from PyTorch (plumbing only).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtdrn_b200.so')

F32, BF16 = 0, 1
F16 = 3            # TDRN_F16: IEEE half activations / weights (MobileNet trunks)


class TdrnError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ('B', 'H', 'W', 'Cin', 'Cout', 'kh', 'kw', 'stride', 'pad', 'dil', 'relu', 'deconv2x2', 'dg',
                 'in_dtype', 'out_dtype')] + [('out_sb', ctypes.c_longlong), ('out_sp', ctypes.c_longlong), ('in_sb', ctypes.c_longlong), ('pool2x2', ctypes.c_int), ('split3', ctypes.c_int), ('split_out', ctypes.c_int)]


class OffsetLevel(ctypes.Structure):
    _fields_ = [('H', ctypes.c_int), ('W', ctypes.c_int), ('prior_off', ctypes.c_int)] + \
               [(n, ctypes.c_void_p) for n in ('w1', 'b1', 'w2', 'b2', 'out1', 'out2', 'out1_nchw')]


class DeformHeadDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ('B', 'H', 'W', 'Cin', 'num_classes', 'dg', 'kh', 'pad', 'kh2', 'pad2', 'P', 'prior_off', 'softmax', 'split')]


class DwPwDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ('B', 'H', 'W', 'Cin', 'Cout', 'stride', 'relu_dw', 'relu_pw')]


_lib = None

EXPORTS = [
    'tdrn_last_error', 'tdrn_version', 'tdrn_launch_count', 'tdrn_prior_box', 'tdrn_deform_conv_forward',
    'tdrn_nms_workspace_bytes', 'tdrn_nms', 'tdrn_nms_host', 'tdrn_nms_rule', 'tdrn_nms_host_rule', 'tdrn_decode', 'tdrn_detect_workspace_bytes',
    'tdrn_detect', 'tdrn_conv2d', 'tdrn_conv2d_tc', 'tdrn_dwconv3x3', 'tdrn_dwconv3x3_io', 'tdrn_conv_dwpw', 'tdrn_conv_first', 'tdrn_conv_stem_pair', 'tdrn_maxpool2x2',
    'tdrn_l2norm', 'tdrn_l2norm_io', 'tdrn_l2norm_pool2x2', 'tdrn_softmax', 'tdrn_nhwc_to_nchw_f32', 'tdrn_nchw_f32_to_nhwc', 'tdrn_split_bf16', 'tdrn_offset_convs', 'tdrn_deform_head', 'tdrn_deform_head_sample', 'tdrn_deform_head_sample_group', 'tdrn_collect_workspace_bytes', 'tdrn_collect_detections', 'tdrn_preprocess', 'tdrn_multiscale_vote_workspace_bytes', 'tdrn_multiscale_vote',
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TdrnError('%s not found: build it with `python -m tdrn_b200.build` (nvcc, sm_100a). '
                            'There is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.tdrn_last_error.restype = ctypes.c_char_p
        L.tdrn_launch_count.restype = ctypes.c_longlong
        L.tdrn_nms_workspace_bytes.restype = ctypes.c_size_t
        L.tdrn_detect_workspace_bytes.restype = ctypes.c_size_t
        L.tdrn_collect_workspace_bytes.restype = ctypes.c_size_t
        L.tdrn_multiscale_vote_workspace_bytes.restype = ctypes.c_size_t
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != 0:
        raise TdrnError('%s failed (rc=%d): %s' % (what or 'tdrn call', rc, lib().tdrn_last_error().decode()))


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_handle():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().tdrn_launch_count())


_probe = None


def probe_lib():
    """libtdrn_probe.so (development probes, csrc/probe/; built by `python -m tdrn_b200.build --probe`)."""
    global _probe
    if _probe is None:
        lib()                                    # the probes link against the product library
        path = os.path.join(os.path.dirname(LIB_PATH), 'libtdrn_probe.so')
        if not os.path.exists(path):
            raise TdrnError('%s not found: build it with `python -m tdrn_b200.build --probe`' % path)
        _probe = ctypes.CDLL(path)
    return _probe
