"""nms mirror (utils/nms_wrapper.py:23-31): ``nms(dets, thresh, force_cpu=False) -> list[int]``.

dets: float32 ndarray [N,5] = (x1,y1,x2,y2,score).  Both branches of the reference run on the GPU here (there is no CPU
implementation by design); ``force_cpu`` only selects the reference branch's suppression RULE:

  * ``force_cpu=True``  -> ``cpu_nms`` (utils/nms/cpu_nms.pyx:17-68): +1 areas, suppress when ``ovr >= thresh``
    (threshold compared as a double) -- the rule ``Detect`` uses;
  * ``force_cpu=False`` -> ``gpu_nms`` (utils/nms/nms_kernel.cu:71): suppress when ``ovr > (float)thresh``.

The two differ only for a pair whose IoU equals the threshold exactly.  Pinned deviation from the reference: boxes with
EQUAL scores are visited lower index first (the reference's ``argsort()[::-1]`` order among ties is an accident of NumPy's
introsort: higher index first for short / already-sorted inputs); tests/test_gpu_postprocess.py pins both facts.
"""
import ctypes

import numpy as np

from .. import _lib


def nms(dets, thresh, force_cpu=False):
    if dets.shape[0] == 0:
        return []
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] != 5:
        raise ValueError('dets must be [N,5]')
    n = dets.shape[0]
    keep = np.empty(n, dtype=np.int32)
    num = ctypes.c_int(0)
    L = _lib.lib()
    _lib.check(L.tdrn_nms_host_rule(keep.ctypes.data_as(ctypes.c_void_p), ctypes.byref(num),
                                    dets.ctypes.data_as(ctypes.c_void_p), n, 5, ctypes.c_double(thresh), -1,
                                    0 if force_cpu else 1),
               'tdrn_nms_host')
    return keep[:num.value].tolist()
