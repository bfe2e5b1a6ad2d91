"""CPU restatement of the pre-processing in front of the hot path.  TEST INFRASTRUCTURE ONLY.

    base_transform(image, size, mean)      reference data/__init__.py:7-12
    img[:, :, (2, 1, 0)] + permute(2,0,1)  reference data/voc0712.py:466-468 (dataset drivers),
                                           test_video_trn.py:89-91 (video: no channel swap)

Parity: PINNED against real OpenCV.  `cv2.resize` lives in OpenCV, a third-party dependency the reference does not pin
(README: "OpenCV"); this image carries OpenCV 4.13.0, and the restatement is bit-identical to it: the reference's own
`base_transform` source executed with that cv2 produced tests/golden/base_transform.npz (oracle/make_golden_preprocess.py),
and tests/test_preprocess.py also compares live with cv2.resize (twelve shapes incl. 1x1 and 5x300 sources, up- and
down-scaling, flipped) wherever cv2 imports.  The resize below restates OpenCV's 8-bit INTER_LINEAR algorithm
(modules/imgproc/src/resize.cpp of OpenCV 3.x / 4.x: `resizeGeneric_` coefficient set-up,
`HResizeLinear<uchar,int,short,2048>`, `VResizeLinear<uchar,int,short, FixedPtCast<int,uchar,22>>`) -- integer
arithmetic, so the result does not depend on the CPU or on SIMD dispatch.
Further cross-checks (tests/test_preprocess.py): identity when sizes match, exact 2x2 box average at scale 2, and agreement
within one grey level with torch's float bilinear interpolation (same half-pixel coordinate mapping).
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS          # INTER_RESIZE_COEF_SCALE = 2048


def _axis(n_dst, n_src, clamp_frac):
    scale = 1.0 / (float(n_dst) / float(n_src))                      # cv::resize: scale = 1. / inv_scale
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)                 # fx = (float)((dx+0.5)*scale_x - 0.5)
    s = np.floor(f).astype(np.int64)                                 # cvFloor
    f = f - s.astype(np.float32)
    if clamp_frac:                                                   # columns: sx < 0 / sx >= width-1 -> fx = 0
        lo = s < 0
        s[lo] = 0; f[lo] = 0
        hi = s >= n_src - 1
        s[hi] = n_src - 1; f[hi] = 0
    a0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)   # saturate_cast<short>: half to even
    a1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)
    return s, a0, a1


def cv2_resize_linear_u8(img, size):
    """img [H,W,C] uint8 -> [size,size,C] uint8, cv2.resize(img, (size, size)) with the default INTER_LINEAR."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    sx, ax0, ax1 = _axis(size, W, True)
    sy, by0, by1 = _axis(size, H, False)
    y0 = np.clip(sy, 0, H - 1)
    y1 = np.clip(sy + 1, 0, H - 1)                                   # rows are clipped, fy is not
    x1 = np.minimum(sx + 1, W - 1)
    src = img.astype(np.int64)
    rows0, rows1 = src[y0], src[y1]                                  # [size, W, C]
    h0 = rows0[:, sx] * ax0[None, :, None] + rows0[:, x1] * ax1[None, :, None]      # HResizeLinear (int)
    h1 = rows1[:, sx] * ax0[None, :, None] + rows1[:, x1] * ax1[None, :, None]
    v = (((by0[:, None, None] * (h0 >> 4)) >> 16) + ((by1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2   # VResizeLinear
    return np.clip(v, 0, 255).astype(np.uint8)


def base_transform(image, size, mean):
    """data/__init__.py:7-12."""
    x = cv2_resize_linear_u8(image, size).astype(np.float32)
    x -= np.asarray(mean, dtype=np.float32)
    return x.astype(np.float32)


def network_input(frames, size, mean, to_rgb, flip=False):
    """[B,H,W,3] uint8 -> [B,3,size,size] float32, as the drivers build it frame by frame (flip: cv2.flip(im, 1) first,
    multi_eval.py:541-544)."""
    out = []
    for f in frames:
        x = base_transform(f[:, ::-1] if flip else f, size, mean)
        if to_rgb:
            x = x[:, :, (2, 1, 0)]                                   # data/voc0712.py:466-467
        out.append(np.transpose(x, (2, 0, 1)))                       # .permute(2, 0, 1)
    return np.stack(out).astype(np.float32)
