"""Result scatter of the evaluation drivers on the device (the step right after the hot path).

The reference walks ``detections[i, j]`` for every image and class, keeps the rows with score > 0, scales the boxes by
the image's (w, h) and copies each piece to the host (evaluate.py:469-483, evaluate_coco.py:140-159).  Here one C-ABI
call compacts the whole batch into rows ``(image, class, x1, y1, x2, y2, score)`` in pixels and one copy brings them back.
"""
import ctypes

import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr, stream_handle


def collect_detections(detections, sizes, max_rows=None):
    """detections [B,C,top_k,5] CUDA fp32 (Detect output); sizes [B,2] = (width, height) per image (tensor / array / list)
    -> (rows [n,7] CUDA fp32 ordered by image, class, rank; n int)."""
    if not detections.is_cuda:
        raise NotImplementedError('detections must be a CUDA tensor: tdrn_b200 has no CPU path')
    det = detections.float().contiguous()
    B, C, K, five = det.shape
    assert five == 5
    wh = torch.as_tensor(np.asarray(sizes, dtype=np.float32) if not torch.is_tensor(sizes) else sizes, dtype=torch.float32)
    wh = wh.reshape(B, 2).to(det.device).contiguous()
    if max_rows is None:
        max_rows = B * (C - 1) * K
    rows = torch.empty(max(max_rows, 1), 7, dtype=torch.float32, device=det.device)
    count = torch.zeros(1, dtype=torch.int32, device=det.device)
    L = _lib.lib()
    nws = L.tdrn_collect_workspace_bytes(B, C)
    ws = torch.empty(max(nws, 4), dtype=torch.uint8, device=det.device)
    check(L.tdrn_collect_detections(ptr(det), ptr(wh), B, C, K, ptr(rows), int(max_rows), ptr(count), ptr(ws),
                                    ctypes.c_size_t(ws.numel()), stream_handle()), 'tdrn_collect_detections')
    n = int(count.item())
    return rows[:min(n, max_rows)], n


def to_all_boxes(rows, num_images, num_classes):
    """rows [n,7] (any device) -> ``all_boxes[cls][img]`` = float32 ndarray [k,5] (x1,y1,x2,y2,score) or [] -- the structure
    evaluate.py:433-434,480-483 builds and write_voc_results_file consumes."""
    r = rows.detach().cpu().numpy()
    all_boxes = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    if len(r):
        img, cls = r[:, 0].astype(np.int64), r[:, 1].astype(np.int64)
        key = img * num_classes + cls
        starts = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
        ends = np.r_[starts[1:], len(r)]
        for s, e in zip(starts, ends):
            all_boxes[cls[s]][img[s]] = np.ascontiguousarray(r[s:e, 2:7], dtype=np.float32)
    return all_boxes
