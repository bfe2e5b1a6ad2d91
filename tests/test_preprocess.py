"""Pre-processing in front of the hot path: the restated cv2 fixed-point resize / base_transform
(oracle/preprocess_ref.py) against independent cross-checks on CPU, and tdrn_preprocess against it on the GPU.
The reference ships no fixture for this step and pins no OpenCV version; the pin is the reference's own `base_transform`
source executed with the OpenCV of this image (4.13.0): tests/golden/base_transform.npz (oracle/make_golden_preprocess.py),
plus live comparisons with cv2 where it is importable.  The kernel and the restatement agree bit for bit."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess_ref as P

MEAN = (104, 117, 123)          # data/config.py: VOC means in cv2 (BGR) channel order


def _frames(b, h, w, seed):
    rng = np.random.RandomState(seed)
    smooth = rng.randint(0, 256, size=(b, h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    up = F.interpolate(torch.from_numpy(smooth).permute(0, 3, 1, 2), size=(h, w), mode='bilinear', align_corners=False)
    img = up.permute(0, 2, 3, 1).numpy() + rng.randn(b, h, w, 3) * 12
    img[:, :3, :3] = 255; img[:, -3:, -3:] = 0                       # saturated corners
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def test_resize_identity_and_exact_halving():
    img = _frames(1, 64, 48, 0)[0]
    sq = img[:48]                                                     # 48x48 -> 48: every coefficient is (2048, 0)
    assert np.array_equal(P.cv2_resize_linear_u8(sq, 48), sq)
    big = _frames(1, 96, 96, 1)[0]
    half = P.cv2_resize_linear_u8(big, 48)                            # scale 2: fx = fy = 0.5 -> (a+b+c+d+2) >> 2
    s = big.astype(np.int64)
    box = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
    assert np.array_equal(half, box.astype(np.uint8))


def test_resize_hand_computed_vector():
    # 2x2 image (two equal rows) -> 4x4.  Columns: fx(dx) = (dx+.5)*.5-.5 = -.25, .25, .75, 1.25 -> clamp, (1536,512),
    # (512,1536), clamp.  Row dy = 1: fy = .25 -> beta = (1536, 512) on two equal source rows; the two products are floored
    # separately, which is where the fixed-point path differs from a float blend (57.5 -> 57, not 58).
    img = np.array([[[10, 20, 30], [200, 100, 50]]], dtype=np.uint8)
    out = P.cv2_resize_linear_u8(np.repeat(img, 2, 0), 4)

    def px(a, b, w0, w1, b0=1536, b1=512):
        h = a * w0 + b * w1
        return (((b0 * (h >> 4)) >> 16) + ((b1 * (h >> 4)) >> 16) + 2) >> 2
    assert out[1, :, 0].tolist() == [px(10, 200, 2048, 0), px(10, 200, 1536, 512), px(10, 200, 512, 1536), px(200, 200, 2048, 0)]
    assert out[1, :, 0].tolist() == [10, 57, 152, 200]
    # h = 10*1536 + 200*512 = 117760; h >> 4 = 7360; (1536*7360) >> 16 = 172, (512*7360) >> 16 = 57; (172+57+2) >> 2 = 57


@pytest.mark.parametrize('h,w,size', [(480, 640, 320), (375, 500, 320), (240, 352, 512), (333, 77, 320), (1080, 1920, 320)])
def test_resize_close_to_float_bilinear(h, w, size):
    img = _frames(1, h, w, h + w)[0]
    got = P.cv2_resize_linear_u8(img, size).astype(np.int32)
    ref = F.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None].double(), size=(size, size), mode='bilinear',
                        align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
    assert np.abs(got - ref).max() <= 1.0 + 1e-9                       # 11-bit coefficients + integer rounding: one grey level


def test_base_transform_and_layout():
    frames = _frames(2, 60, 80, 5)
    x = P.network_input(frames, 32, MEAN, to_rgb=True)
    assert x.shape == (2, 3, 32, 32) and x.dtype == np.float32
    r = P.cv2_resize_linear_u8(frames[1], 32).astype(np.float32)
    assert np.array_equal(x[1, 0], r[:, :, 2] - 123) and np.array_equal(x[1, 2], r[:, :, 0] - 104)
    y = P.network_input(frames, 32, MEAN, to_rgb=False)
    assert np.array_equal(y[1, 0], r[:, :, 0] - 104)


def _golden_cases(golden):
    g = golden('base_transform')
    mean = g['mean']
    i = 0
    while 'in_%d' % i in g.files:
        yield g['in_%d' % i], g['res_%d' % i].astype(np.float32) - mean, mean
        i += 1


def test_restatement_matches_the_reference_base_transform_golden(golden):
    """oracle/preprocess_ref.base_transform == the reference's base_transform (data/__init__.py:7-12) run with real cv2."""
    n = 0
    for img, ref, mean in _golden_cases(golden):
        got = P.base_transform(img, ref.shape[0], mean)
        assert got.dtype == np.float32 and np.array_equal(got, ref)
        n += 1
    assert n == 8


@pytest.mark.parametrize('h,w,size', [(480, 640, 320), (375, 500, 320), (240, 352, 512), (333, 77, 320), (1080, 1920, 320),
                                      (320, 320, 320), (100, 100, 704), (720, 1280, 192), (321, 319, 320), (1, 1, 8), (5, 300, 17)])
def test_restatement_matches_cv2_live(h, w, size):
    """Bit-exact against cv2.resize (default INTER_LINEAR) of whatever OpenCV is installed; skipped where there is none."""
    cv2 = pytest.importorskip('cv2')
    img = _frames(1, max(h, 16), max(w, 16), 3 * h + w)[0, :h, :w].copy()
    assert np.array_equal(P.cv2_resize_linear_u8(img, size), cv2.resize(img, (size, size)).reshape(size, size, 3))
    flipped = cv2.flip(img, 1).reshape(img.shape)                     # multi_eval.py:541-544
    assert np.array_equal(P.cv2_resize_linear_u8(img[:, ::-1], size), cv2.resize(flipped, (size, size)).reshape(size, size, 3))


def test_reference_base_transform_source_live():
    """Build container only: the reference's own function source + cv2 vs the restatement, on frames no fixture holds."""
    import os
    pytest.importorskip('cv2')
    from oracle import make_golden_preprocess as G
    if not os.path.exists(G.REF):
        pytest.skip('reference checkout not present (GPU box)')
    fn = G.reference_base_transform()
    for seed, (h, w, size) in enumerate([(480, 640, 320), (375, 500, 512), (300, 300, 300)]):
        img = _frames(1, h, w, 70 + seed)[0]
        assert np.array_equal(fn(img.copy(), size, np.array(MEAN, dtype=np.float32)), P.base_transform(img, size, MEAN))


@pytest.mark.gpu
def test_preprocess_kernel_matches_the_reference_base_transform_golden(golden):
    """tdrn_preprocess == the reference's base_transform run with real cv2 (committed fixture), bit for bit."""
    from tdrn_b200.data import base_transform
    for img, ref, mean in _golden_cases(golden):
        got = base_transform(img, ref.shape[0], tuple(float(m) for m in mean))
        assert got.shape == ref.shape and got.dtype == np.float32 and np.array_equal(got, ref)

@pytest.mark.gpu
@pytest.mark.parametrize('flip', [False, True])
@pytest.mark.parametrize('b,h,w,size,rgb', [(3, 480, 640, 320, False), (2, 375, 500, 320, True), (1, 240, 352, 512, True),
                                            (2, 333, 77, 320, False), (1, 1080, 1920, 320, False), (4, 320, 320, 320, True),
                                            (1, 1, 1, 8, False), (2, 5, 300, 17, True)])
def test_preprocess_kernel_bit_exact(b, h, w, size, rgb, flip):
    from tdrn_b200 import ops
    frames = _frames(b, max(h, 16), max(w, 16), b * h + w)[:, :h, :w].copy()
    ref = P.network_input(frames, size, MEAN, to_rgb=rgb, flip=flip)
    out = ops.preprocess(torch.from_numpy(frames).cuda(), size, MEAN, swap_rb=rgb, flip_lr=flip)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_base_transform_mirror_and_errors():
    from tdrn_b200.data import base_transform, BaseTransform, preprocess_frames
    from tdrn_b200 import _lib
    img = _frames(1, 123, 211, 9)[0]
    ref = P.base_transform(img, 320, MEAN)
    got = base_transform(img, 320, MEAN)
    assert got.shape == (320, 320, 3) and got.dtype == np.float32 and np.array_equal(got, ref)
    t, bx, lb = BaseTransform(320, MEAN)(img, 'boxes', 'labels')       # data/__init__.py:19-23: boxes / labels pass through
    assert np.array_equal(t, ref) and bx == 'boxes' and lb == 'labels'
    x = preprocess_frames(img, 320, MEAN, to_rgb=True)
    assert tuple(x.shape) == (1, 3, 320, 320) and x.is_cuda
    with pytest.raises(TypeError):
        preprocess_frames(img.astype(np.float32), 320, MEAN)
    with pytest.raises(_lib.TdrnError):
        preprocess_frames(img, 0, MEAN)
