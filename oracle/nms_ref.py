"""NumPy restatement of the reference CPU NMS.  TEST INFRASTRUCTURE ONLY.

Follows utils/nms/cpu_nms.pyx:17-68 (hard NMS, what Detect runs: layers/functions/detection.py:60
calls utils/nms_wrapper.py:23-31 with force_cpu=True -> cpu_nms).

Pinned tie rule: the reference orders candidates with
``scores.argsort()[::-1]`` (cpu_nms.pyx:25), whose order among *equal* scores is an accident of
NumPy's introsort.  Both this oracle and the CUDA kernel use "descending score, ties -> lower
original index first", i.e. ``np.argsort(-scores, kind='stable')``.  For distinct scores this is
identical to the reference.

Parity status: pinned against the reference's own compiled Cython module (oracle/build_ref_nms.py builds
cpu_nms.pyx through a two-token dtype respelling; tests/test_ref_cython_nms.py compares keep lists, the
pair-exactly-on-the-threshold case included) and cross-checked against the reference's importable
utils/nms/py_cpu_nms.py (identical except at ovr == thresh exactly).
"""
import numpy as np


def order_desc_stable(scores):
    return np.argsort(-scores.astype(np.float32), kind="stable")


def cpu_nms(dets, thresh, max_keep=None):
    """dets float32 [N,5] = (x1,y1,x2,y2,score); returns list[int] of kept indices, score order.

    ``max_keep``: stop after that many boxes are kept.  Detect only consumes keep[:top_k]
    (detection.py:61-63) and greedy NMS visits candidates in descending score order, so the
    truncated list equals the first ``max_keep`` entries of the full list.
    """
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return []                                                   # nms_wrapper.py:26-27
    x1, y1, x2, y2, scores = (dets[:, k] for k in range(5))         # pyx:18-22
    one = np.float32(1)
    areas = (x2 - x1 + one) * (y2 - y1 + one)                       # pyx:24 (float32)
    order = order_desc_stable(scores)                               # pyx:25 + pinned tie rule
    suppressed = np.zeros(n, dtype=bool)                            # pyx:28-29
    thresh = float(thresh)                                          # `np.float thresh` is a C double
    keep = []
    for _i in range(n):                                             # pyx:43
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        if max_keep is not None and len(keep) >= max_keep:
            break
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])                           # pyx:57-60
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1 + one)              # pyx:61
        h = np.maximum(np.float32(0), yy2 - yy1 + one)              # pyx:62
        inter = w * h                                               # pyx:63 float32
        ovr = inter / (areas[i] + areas[rest] - inter)              # pyx:64 float32 divide
        # pyx:65: float32 ovr compared with the double thresh
        suppressed[rest[ovr.astype(np.float64) >= thresh]] = True
    return keep
