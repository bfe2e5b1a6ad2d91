"""Mirror of the pieces of the reference's `data` package that sit on the hot path's boundary: the prior-box
dictionaries (data/config.py) and base_transform / BaseTransform (data/__init__.py:7-23), computed on the device."""
import numpy as np
import torch

from .config import mb_cfg, multi_cfg, multi_cfg_512, multi_scale, VOC_320, VOC_512_RefineDet
from .. import ops


def preprocess_frames(frames, size, mean, to_rgb=False, device='cuda', flip=False):
    """Batch form used by a serving loop: uint8 frames [B,H,W,3] (numpy or torch, cv2 BGR order) -> the network input
    x [B,3,size,size] fp32 on the device, i.e. what the reference builds with base_transform(frame, size, mean),
    (optionally `img[:, :, (2, 1, 0)]`, data/voc0712.py:466-467) and `.permute(2, 0, 1)` per frame."""
    f = torch.as_tensor(np.ascontiguousarray(frames) if isinstance(frames, np.ndarray) else frames)
    if f.dim() == 3:
        f = f.unsqueeze(0)
    return ops.preprocess(f.to(device, non_blocking=True), size, mean, swap_rb=to_rgb, flip_lr=flip)


def base_transform(image, size, mean):
    """data/__init__.py:7-12: HWC uint8 image -> HWC float32 numpy array (resized, mean-subtracted).  Same return type
    as the reference (host array); callers that stay on the device should use preprocess_frames instead."""
    x = preprocess_frames(image, size, np.asarray(mean, dtype=np.float32).reshape(-1)[:3])
    return x[0].permute(1, 2, 0).contiguous().cpu().numpy()


class BaseTransform(object):
    """data/__init__.py:14-23."""

    def __init__(self, size, mean):
        self.size = size
        self.mean = np.array(mean, dtype=np.float32)

    def __call__(self, image, boxes=None, labels=None):
        return base_transform(image, self.size, self.mean), boxes, labels
