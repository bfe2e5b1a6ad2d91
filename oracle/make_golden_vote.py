"""Generate tests/golden/bbox_vote.npz by executing the REAL `bbox_vote` of the reference: its source is taken from
/root/reference/multi_eval.py:453-494 with `ast` (the module itself cannot be imported: it parses command-line arguments,
imports cv2 and opens datasets at import time) and run, unmodified, on seeded detections.

    python -m oracle.make_golden_vote        (build container only; /root/reference must exist)

TEST INFRASTRUCTURE ONLY.  The fixture pins oracle/multi_scale_ref.bbox_vote (CPU test) and, through it, the device
kernel tdrn_multiscale_vote (GPU test).  Scores are distinct, so the reference's unstable `argsort()[::-1]` has one
possible order.  NumPy version at generation time is stored in the fixture.
"""
import ast
import os

import numpy as np

REF = '/root/reference/multi_eval.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'bbox_vote.npz')


def reference_bbox_vote():
    tree = ast.parse(open(REF).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'bbox_vote']
    assert len(fn) == 1
    ns = {'np': np}
    exec(compile(ast.Module(fn, []), REF, 'exec'), ns)
    return ns['bbox_vote']


def make_case(seed, n, clusters):
    rng = np.random.RandomState(seed)
    ctr = rng.rand(clusters, 2).astype(np.float32) * 400 + 50
    size = rng.rand(clusters, 2).astype(np.float32) * 120 + 10
    k = rng.randint(0, clusters, n)
    c = ctr[k] + rng.randn(n, 2).astype(np.float32) * 3
    s = size[k] * (1 + rng.randn(n, 2).astype(np.float32) * 0.06)
    score = rng.permutation(n).astype(np.float32) / np.float32(n) * np.float32(0.98) + np.float32(0.01)   # distinct
    return np.hstack((c - s / 2, c + s / 2, score[:, None])).astype(np.float32)


def main():
    vote = reference_bbox_vote()
    rec = {'numpy_version': np.array(np.__version__)}
    for i, (n, clusters) in enumerate([(2, 1), (9, 3), (40, 6), (137, 10), (300, 4), (1000, 25), (2800, 40)]):
        det = make_case(100 + i, n, clusters)
        out = np.asarray(vote(det.copy()))
        rec['in_%d' % i] = det
        rec['out_%d' % i] = out.astype(np.float64)
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, {k: v.shape for k, v in rec.items() if k.startswith('out')})


if __name__ == '__main__':
    main()
