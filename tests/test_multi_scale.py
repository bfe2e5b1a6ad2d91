"""Multi-scale / flip testing merge: the NumPy restatement of multi_eval.py:453-494,557-640 on CPU,
tdrn_multiscale_vote and the MultiScaleTester driver against it on the GPU (bit-exact: float32 arithmetic in NumPy's order)."""
import numpy as np
import pytest
import torch

from oracle import multi_scale_ref as M


def test_numpy_summation_orders_are_what_the_kernel_emulates():
    rng = np.random.RandomState(0)
    for n in list(range(1, 40)) + [64, 100, 127, 128, 129, 130, 137, 200, 255, 256, 257, 300, 1000, 2800]:
        for _ in range(3):
            acc = rng.rand(n, 5).astype(np.float32)
            assert np.sum(acc[:, -1:]) == M.np_pairwise_sum_f32(acc[:, 4])          # multi_eval.py:486 denominator
            seq = acc[0, 0:4].copy()
            for i in range(1, n):
                seq = (seq + acc[i, 0:4]).astype(np.float32)
            assert np.array_equal(np.sum(acc[:, 0:4], axis=0), seq)                  # :486 numerator: row after row


def test_bbox_vote_hand_case():
    det = np.array([[10, 10, 50, 50, 0.9],          # head
                    [12, 12, 52, 52, 0.6],          # IoU with head 0.82 -> merged
                    [200, 200, 240, 260, 0.8],      # far away -> own group
                    [11, 9, 49, 51, 0.3]], np.float32)   # merged with head too
    out = M.bbox_vote(det)
    assert out.shape == (2, 5) and out.dtype == np.float32
    wsum = np.float32(0.9) + np.float32(0.6) + np.float32(0.3)
    x1 = (np.float32(10) * np.float32(0.9) + np.float32(12) * np.float32(0.6) + np.float32(11) * np.float32(0.3)) / wsum
    assert out[0, 4] == np.float32(0.9) and abs(out[0, 0] - x1) < 1e-5
    assert np.array_equal(out[1], det[2])
    assert np.array_equal(M.bbox_vote(det[:1]), det[:1])                             # <= 1 row: returned unchanged (:454-455)


def test_gather_unflips_scales_and_filters():
    C, top_k = 3, 4
    d = np.zeros((C, top_k, 5), np.float32)
    d[1, 0] = [0.9, 0.1, 0.2, 0.3, 0.6]             # 100 x 150 px at 500 x 375
    d[1, 1] = [0.5, 0.50, 0.50, 0.52, 0.53]         # 11 x 12.25 px (+1): below the 32 px rule of scale 192
    got = M.gather_class([(192, False, d), (192, True, d)], 1, 500, 375, 320)
    assert got.shape == (2, 5)                       # the small box is dropped in both passes (longer side <= 32)
    assert got[0, 0] == np.float32(0.1) * np.float32(500) and got[0, 2] == np.float32(0.3) * np.float32(500)
    assert got[1, 0] == (np.float32(1) - np.float32(0.3)) * np.float32(500)         # flipped pass: x1' = (1 - x2) * w
    assert got[1, 2] == (np.float32(1) - np.float32(0.1)) * np.float32(500)
    assert M.gather_class([(192, False, d)], 2, 500, 375, 320).shape == (0, 5)      # class without detections


def _synthetic_passes(scales, C, top_k, seed):
    rng = np.random.RandomState(seed)
    centers = rng.rand(C, 12, 2).astype(np.float32) * 0.8 + 0.1
    sizes = (rng.rand(C, 12, 2).astype(np.float32) * 0.3 + 0.02)
    passes = []
    for v in scales:
        for flip in (False, True):
            det = np.zeros((C, top_k, 5), np.float32)
            for c in range(1, C):
                n = int(rng.randint(0, top_k + 1)) if c % 5 else 0                   # some classes / passes empty
                if n == 0:
                    continue
                k = rng.randint(0, 12, n)
                ctr = centers[c, k] + rng.randn(n, 2).astype(np.float32) * 0.01
                wh = sizes[c, k] * (1 + rng.randn(n, 2).astype(np.float32) * 0.05)
                if flip:
                    ctr[:, 0] = 1 - ctr[:, 0]
                sc = np.sort(rng.rand(n).astype(np.float32) * 0.99 + 0.01)[::-1]
                det[c, :n, 0] = sc
                det[c, :n, 1:3] = ctr - wh / 2
                det[c, :n, 3:5] = ctr + wh / 2
            passes.append((v, flip, det))
    return passes


@pytest.mark.gpu
@pytest.mark.parametrize('base,scales,C,top_k,wh', [(320, [192, 320, 384, 448, 512, 576, 704], 21, 200, (500, 375)),
                                                    (512, [320, 512, 640, 1216], 9, 300, (640, 480)),
                                                    (320, [320], 4, 7, (33, 21))])
def test_multiscale_vote_kernel_bit_exact(base, scales, C, top_k, wh):
    from tdrn_b200 import ops
    from tdrn_b200.utils.multi_scale import SIZE_RULES
    passes = _synthetic_passes(scales, C, top_k, seed=base + C)
    w, h = wh
    ref = M.multi_scale_merge(passes, C, w, h, base)
    dets = torch.from_numpy(np.stack([p[2] for p in passes])).cuda()
    rules = [SIZE_RULES[(base, p[0])] for p in passes]
    rows, cnt = ops.multiscale_vote(dets, [p[1] for p in passes], [r[0] for r in rules], [r[1] for r in rules], w, h)
    rows, cnt = rows.cpu().numpy(), cnt.cpu().numpy()
    assert cnt[0] == 0
    merged = 0
    for j in range(1, C):
        assert cnt[j] == ref[j].shape[0], (j, cnt[j], ref[j].shape)
        assert np.array_equal(rows[j, :cnt[j]], ref[j]), j
        merged += sum(p[2][j, :, 0].astype(bool).sum() for p in passes) - cnt[j]
    assert merged > 0                                                                # the case does exercise merging


@pytest.mark.gpu
def test_multiscale_tester_end_to_end():
    from tdrn_b200.model.dualrefinedet_vggbn import build_net
    from tdrn_b200.layers.functions import Detect
    from tdrn_b200.utils.multi_scale import MultiScaleTester
    from tdrn_b200.utils.synthetic import randomize_
    torch.manual_seed(0)
    net = build_net('test', 320, 21, 1024, 1, True, True)
    randomize_(net, seed=0)
    net = net.cuda().eval()
    det = Detect(21, 0, 50, 0.01, 0.45)
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(150, 200, 3)).astype(np.uint8)
    tester = MultiScaleTester(net, det, 320, (104, 117, 123), scales=[192, 320, 384])
    dets, flips, rules, thrs = tester.passes(img)
    assert tuple(dets.shape) == (6, 21, 50, 5)
    got = tester.detect(img)
    passes = [(v, f, dets[i].cpu().numpy()) for i, (v, f) in enumerate((v, f) for v in (192, 320, 384) for f in (False, True))]
    ref = M.multi_scale_merge(passes, 21, 200, 150, 320)
    assert len(got) == 21 and got[0].shape[0] == 0
    for j in range(1, 21):
        assert np.array_equal(got[j], ref[j]), j
    with pytest.raises(ValueError):
        MultiScaleTester(net, det, 320, (104, 117, 123), scales=[300])


def _vote_golden(golden):
    g = golden('bbox_vote')
    i = 0
    while 'in_%d' % i in g:
        yield g['in_%d' % i], g['out_%d' % i]
        i += 1


def test_bbox_vote_restatement_matches_the_reference_function(golden):
    """tests/golden/bbox_vote.npz holds outputs of the reference's OWN bbox_vote (multi_eval.py:453-494, executed from its
    source by oracle/make_golden_vote.py) on seeded detections, up to 2 800 boxes with ~50-member groups."""
    n_cases = 0
    for det, ref in _vote_golden(golden):
        got = M.bbox_vote(det)
        assert got.shape == ref.shape and np.array_equal(got.astype(np.float64), ref)
        n_cases += 1
    assert n_cases == 7


@pytest.mark.gpu
def test_multiscale_vote_kernel_matches_the_reference_function(golden):
    """The device kernel against the same fixture: one pass, one class, identity scaling, a size rule that keeps everything."""
    from tdrn_b200 import ops
    for det, ref in _vote_golden(golden):
        n = det.shape[0]
        d = np.zeros((1, 2, n, 5), np.float32)
        d[0, 1, :, 0] = det[:, 4]
        d[0, 1, :, 1:] = det[:, :4]
        rows, cnt = ops.multiscale_vote(torch.from_numpy(d).cuda(), [False], [0], [-1.0], 1, 1)
        rows, cnt = rows.cpu().numpy(), cnt.cpu().numpy()
        assert cnt[1] == ref.shape[0]
        assert np.array_equal(rows[1, :cnt[1]].astype(np.float64), ref)


@pytest.mark.skipif(not __import__('os').path.exists('/root/reference/data/config.py'),
                    reason='reference checkout only exists in the build container')
def test_multi_scale_prior_dicts_equal_the_reference_values():
    """multi_cfg / multi_cfg_512 / multi_scale against the dict literals of data/config.py:139-261 and multi_eval.py:21-24
    (read with `ast`: importing the reference's data package opens dataset files)."""
    import ast
    from tdrn_b200.data import config as C
    ns = {}
    for path, want in (('/root/reference/data/config.py', None), ('/root/reference/multi_eval.py', {'multi_scale'})):
        for node in ast.parse(open(path).read()).body:
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict):
                name = node.targets[0].id if isinstance(node.targets[0], ast.Name) else None
                if want is not None and name not in want:
                    continue
                try:
                    exec(compile(ast.Module([node], []), path, 'exec'), ns)
                except NameError:
                    pass                                      # dicts that reference dataset constants: not needed here
    assert ns['multi_cfg'] == C.multi_cfg and ns['multi_cfg_512'] == C.multi_cfg_512
    assert ns['multi_scale'] == C.multi_scale
    assert ns['VOC_320'] == C.VOC_320 and ns['VOC_512_RefineDet'] == C.VOC_512_RefineDet
    assert ns['mb_cfg'] == C.mb_cfg and list(ns['mb_cfg']) == list(C.mb_cfg)      # data/config.py:257-258, every entry
