"""ctypes bindings of oracle/_build/liboracle.so (scalar C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None
_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path):
            path = _build.build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_deform_conv_forward.restype = ctypes.c_int
        _lib.oracle_cpu_nms.restype = ctypes.c_int
        _lib.oracle_cpu_nms.argtypes = [_f, ctypes.c_int, ctypes.c_double, ctypes.c_int, _i]
        _lib.oracle_detect.restype = ctypes.c_int
        _lib.oracle_detect.argtypes = [_f, _f, _f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_double, _f]
        _lib.oracle_prior_box.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def deform_conv_forward(inp, off, weight, stride=1, pad=0, dil=1, dg=1):
    inp, off, weight = _f32(inp), _f32(off), _f32(weight)
    b, c, h, w = inp.shape
    cout, _, kh, kw = weight.shape
    ho = (h + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    wo = (w + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    out = np.empty((b, cout, ho, wo), np.float32)
    rc = lib().oracle_deform_conv_forward(_fp(inp), _fp(off), _fp(weight), _fp(out), b, c, h, w, cout,
                                          kh, kw, stride, pad, dil, dg)
    if rc != 0:
        raise RuntimeError('oracle_deform_conv_forward rc=%d' % rc)
    return out


def decode(loc, priors, arm_loc=None):
    loc, priors = _f32(loc), _f32(priors)
    b, p, _ = loc.shape
    out = np.empty((b, p, 4), np.float32)
    arm = _f32(arm_loc) if arm_loc is not None else None
    lib().oracle_decode(_fp(loc), _fp(priors), _fp(arm) if arm is not None else None, b, p, _fp(out))
    return out


def cpu_nms(dets, thresh, max_keep=0):
    dets = _f32(dets)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), np.int32)
    nk = lib().oracle_cpu_nms(_fp(dets), n, float(thresh), int(max_keep), keep.ctypes.data_as(_i))
    return keep[:nk].tolist()


def detect(boxes, conf, scale, num_classes, top_k, conf_thresh, nms_thresh):
    boxes, conf, scale = _f32(boxes), _f32(conf), _f32(scale)
    b, p, _ = boxes.shape
    out = np.empty((b, num_classes, top_k, 5), np.float32)
    rc = lib().oracle_detect(_fp(boxes), _fp(conf), _fp(scale), b, p, num_classes, top_k,
                             float(np.float32(conf_thresh)), float(nms_thresh), _fp(out))
    if rc != 0:
        raise RuntimeError('oracle_detect rc=%d' % rc)
    return out


def prior_box(cfg):
    n = len(cfg['feature_maps'])
    ia = lambda v: (ctypes.c_int * len(v))(*v)
    da = lambda v: (ctypes.c_double * len(v))(*[float(t) for t in v])
    ars = []
    for a in cfg['aspect_ratios']:
        ars += list(a) + [0] * (4 - len(a))
    n_ar = [len(a) for a in cfg['aspect_ratios']]
    mx = da(cfg['max_sizes']) if len(cfg['max_sizes']) else None
    args = (cfg['min_dim'], n, ia(cfg['feature_maps']), ia(cfg['steps']), da(cfg['min_sizes']), mx,
            ia(n_ar), da(ars), int(bool(cfg['flip'])), int(bool(cfg['clip'])))
    p = lib().oracle_prior_box(*args, None)
    out = np.empty((p, 4), np.float32)
    lib().oracle_prior_box(*args, _fp(out))
    return out
