"""Generate tests/golden/base_transform.npz by executing the REAL `base_transform` of the reference (its source is taken
from /root/reference/data/__init__.py:7-12 with `ast`; the package itself cannot be imported: data/coco.py opens a label
file at import time) with the REAL cv2 of this image on seeded frames.

    python -m oracle.make_golden_preprocess        (build container only; /root/reference and cv2 must exist)

TEST INFRASTRUCTURE ONLY.  The fixture pins oracle/preprocess_ref.py (CPU test) and, through the same arrays, the device
kernel tdrn_preprocess (GPU test).  The reference pins no OpenCV version (README: "OpenCV"); the version that produced the
fixture is stored in it.  cv2's 8-bit INTER_LINEAR path is integer arithmetic, so the result does not depend on the CPU.
"""
import ast
import os

import numpy as np

REF = '/root/reference/data/__init__.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'base_transform.npz')
MEAN = (104, 117, 123)
CASES = [(120, 160, 96), (75, 100, 64), (37, 53, 80), (120, 176, 128), (64, 64, 64), (133, 77, 96), (50, 70, 192), (9, 300, 40)]


def reference_base_transform():
    import cv2
    tree = ast.parse(open(REF).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'base_transform']
    assert len(fn) == 1
    ns = {'np': np, 'cv2': cv2}
    exec(compile(ast.Module(fn, []), REF, 'exec'), ns)
    return ns['base_transform']


def make_frame(h, w, seed):
    """Smooth structure + noise + saturated corners (so that rounding, clamping and both borders are exercised)."""
    rng = np.random.RandomState(seed)
    gy, gx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([127 + 120 * np.sin(gx / (3.0 + c) + c) * np.cos(gy / (5.0 + c)) for c in range(3)], -1)
    img = base + rng.randn(h, w, 3) * 20
    img[:2, :2] = 255; img[-2:, -2:] = 0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def main():
    import cv2
    fn = reference_base_transform()
    rec = {'cv2_version': np.array(cv2.__version__), 'mean': np.asarray(MEAN, np.float32)}
    for i, (h, w, size) in enumerate(CASES):
        img = make_frame(h, w, 40 + i)
        out = fn(img.copy(), size, np.array(MEAN, dtype=np.float32))          # BaseTransform passes the mean as float32 (:17)
        assert out.dtype == np.float32 and out.shape == (size, size, 3)
        # stored as the resized uint8 image: the reference's output is that image as float32 minus the float32 mean, exactly
        res = np.rint(out + np.asarray(MEAN, np.float32)).astype(np.uint8)
        assert np.array_equal(res.astype(np.float32) - np.asarray(MEAN, np.float32), out)
        rec['in_%d' % i] = img
        rec['res_%d' % i] = res
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, cv2.__version__, {k: v.shape for k, v in rec.items() if k.startswith('res')})


if __name__ == '__main__':
    main()
