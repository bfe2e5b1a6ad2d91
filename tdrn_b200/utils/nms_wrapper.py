"""nms mirror (utils/nms_wrapper.py:23-31): ``nms(dets, thresh, force_cpu=False) -> list[int]``.

dets: float32 ndarray [N,5] = (x1,y1,x2,y2,score).  Runs the CUDA NMS (tdrn_nms_host) with the CPU
rule Detect uses (cpu_nms.pyx: +1 areas, suppress when ovr >= thresh).  ``force_cpu`` is accepted for
signature compatibility and ignored: there is no CPU implementation by design.
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
    _lib.check(L.tdrn_nms_host(keep.ctypes.data_as(ctypes.c_void_p), ctypes.byref(num),
                               dets.ctypes.data_as(ctypes.c_void_p), n, 5, ctypes.c_double(thresh), -1),
               'tdrn_nms_host')
    return keep[:num.value].tolist()
