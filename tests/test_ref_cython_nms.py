"""The reference's OWN Cython CPU NMS (utils/nms/cpu_nms.pyx, what Detect calls through utils/nms_wrapper.py:23-31), built
by oracle/build_ref_nms.py, against the NumPy and C restatements in oracle/ -- CPU tests.  This is the pin for the one
operator DESIGN.md used to list as "parity unpinned": the `ovr >= thresh` comparison (cpu_nms.pyx:65; the reference's GPU
kernel and py_cpu_nms.py use `>`), the `+1` areas and the float32 IoU arithmetic.

Scores are distinct in every case: among equal scores the reference's order (`scores.argsort()[::-1]`, cpu_nms.pyx:25) is an
accident of NumPy's introsort, and both the oracle and the product pin "ties -> lower index first" instead."""
import numpy as np
import pytest

from oracle import build_ref_nms, nms_ref, c_oracle


@pytest.fixture(scope='module')
def ref_nms():
    try:
        build_ref_nms.build()
    except Exception as e:                                         # noqa: BLE001
        pytest.skip('reference cpu_nms.pyx could not be built here: %r' % (e,))
    mod = build_ref_nms.load()
    if mod is None:
        pytest.skip('oracle/_ref/cpu_nms*.so not present (needs /root/reference + Cython to build)')
    return mod.cpu_nms


def boxes(n, seed, span=300.0, size=80.0):
    rng = np.random.RandomState(seed)
    xy = rng.rand(n, 2) * span
    wh = rng.rand(n, 2) * size + 4
    s = rng.permutation(n).astype(np.float32) / n + rng.rand(n).astype(np.float32) * (0.5 / n)     # distinct scores
    assert len(np.unique(s)) == n
    return np.hstack([xy, xy + wh, s[:, None]]).astype(np.float32)


@pytest.mark.parametrize('n,seed,thresh,span', [(1, 0, 0.45, 300), (2, 1, 0.45, 50), (33, 2, 0.3, 100), (257, 3, 0.45, 300),
                                                (1000, 4, 0.45, 300), (1000, 5, 0.1, 120), (3000, 6, 0.7, 200),
                                                (6375, 7, 0.45, 320)])
def test_restatements_equal_the_reference_cython_nms(ref_nms, n, seed, thresh, span):
    d = boxes(n, seed, span)
    ref = [int(i) for i in ref_nms(d, thresh)]
    assert nms_ref.cpu_nms(d, thresh) == ref
    assert [int(i) for i in c_oracle.cpu_nms(d, thresh)] == ref
    for k in (1, 5, 200):                                       # the early exit only truncates (detection.py:61-63)
        assert nms_ref.cpu_nms(d, thresh, max_keep=k) == ref[:k]
        assert [int(i) for i in c_oracle.cpu_nms(d, thresh, k)] == ref[:k]


def test_pair_exactly_on_the_threshold_is_suppressed(ref_nms):
    """IoU == thresh exactly: cpu_nms.pyx:65 suppresses (`>=`), unlike the reference's own GPU kernel (nms_kernel.cu:71, `>`)
    and py_cpu_nms.py:35.  Boxes (0,0,9,9) and (0,0,9,19): areas 100 and 200 with the +1 convention, intersection 100,
    IoU = 100 / 200 = 0.5, exact in float32."""
    d = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 130, 0.7]], np.float32)
    assert [int(i) for i in ref_nms(d, 0.5)] == [0, 2]
    assert nms_ref.cpu_nms(d, 0.5) == [0, 2]
    assert [int(i) for i in c_oracle.cpu_nms(d, 0.5)] == [0, 2]
    # a hair above the pair's IoU and it survives
    assert [int(i) for i in ref_nms(d, 0.5000001)] == [0, 1, 2] == nms_ref.cpu_nms(d, 0.5000001)


def test_thresholds_between_float32_and_double(ref_nms):
    """The float32 IoU is compared with the DOUBLE thresh (`np.float thresh`): 0.45 as a double is not a float32 value; an
    IoU that rounds to float32(0.45) sits above the double 0.45 or below it, and the restatements must agree with the
    compiled code on which.  Dense grid of near-threshold pairs."""
    rng = np.random.RandomState(9)
    rows = []
    for i in range(400):
        w = 20 + i % 37
        h = 10 + i % 23
        # second box shares the corner; its height is tuned so that the IoU lands near 0.45
        h2 = int(round(h / 0.45)) + (i % 3) - 1
        x0, y0 = 500.0 * i, 0.0
        rows.append([x0, y0, x0 + w - 1, y0 + h - 1, 0.9 - 1e-4 * i])
        rows.append([x0, y0, x0 + w - 1, y0 + h2 - 1, 0.8 - 1e-4 * i - 1e-5])
    d = np.asarray(rows, np.float32)
    for t in (0.45, float(np.float32(0.45)), 0.44, 0.46):
        ref = [int(i) for i in ref_nms(d, t)]
        assert nms_ref.cpu_nms(d, t) == ref
        assert [int(i) for i in c_oracle.cpu_nms(d, t)] == ref
