"""Restatement of the per-image result scatter of the reference's evaluation loop (evaluate.py:469-483; the same lines
appear in evaluate_coco.py:140-159 and evaluate_trn.py).  TEST INFRASTRUCTURE ONLY.

    for j in 1..C-1:  dets = detections[0, j]; skip if dets.sum() == 0
        keep rows with score > 0; boxes[:, 0::2] *= w; boxes[:, 1::2] *= h   (fp32, in place)
        all_boxes[j][i] = hstack(boxes, scores[:, None]).astype(float32)
"""
import numpy as np
import torch


def all_boxes_ref(detections, sizes):
    """detections [B,C,top_k,5] CPU fp32 tensor, sizes [(w, h)] per image -> all_boxes[cls][img] like evaluate.py."""
    det = detections.detach().cpu().float()
    B, C = det.shape[0], det.shape[1]
    all_boxes = [[[] for _ in range(B)] for _ in range(C)]
    for i in range(B):
        w, h = sizes[i]
        for j in range(1, C):                                        # evaluate.py:469 (skip background)
            dets = det[i, j, :].clone()
            if dets.sum() == 0:                                      # :471
                continue
            mask = dets[:, 0].gt(0.).expand(dets.size(-1), dets.size(0)).t()      # :473
            dets = torch.masked_select(dets, mask).view(-1, dets.size(-1))        # :474
            boxes = dets[:, 1:]                                      # :475
            boxes[:, 0] *= w                                         # :476-479
            boxes[:, 2] *= w
            boxes[:, 1] *= h
            boxes[:, 3] *= h
            scores = dets[:, 0].numpy()
            all_boxes[j][i] = np.hstack((boxes.numpy(), scores[:, np.newaxis])).astype(np.float32, copy=False)   # :481-483
    return all_boxes
