"""GPU parity: decode / NMS / Detect through the C ABI.  Integer outputs are compared bit-exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand_dets(n, seed, spread=300.0, size=80.0):
    rng = np.random.RandomState(seed)
    xy = rng.rand(n, 2) * spread
    wh = rng.rand(n, 2) * size + 4
    return np.hstack([xy, xy + wh, rng.rand(n, 1)]).astype(np.float32)


@pytest.mark.parametrize('n,seed,thresh', [(1, 0, 0.45), (2, 1, 0.45), (33, 2, 0.3), (256, 3, 0.45), (257, 4, 0.45),
                                           (1000, 5, 0.45), (6375, 6, 0.45), (6375, 7, 0.7), (16320, 8, 0.45)])
def test_nms_wrapper_bit_exact(n, seed, thresh):
    from oracle import c_oracle as C
    from tdrn_b200.utils.nms_wrapper import nms
    dets = _rand_dets(n, seed, spread=300.0 if n < 5000 else 900.0)
    keep = nms(dets, thresh, force_cpu=True)
    assert keep == C.cpu_nms(dets, thresh)


def test_nms_dense_overlaps_ties_and_max_keep(golden):
    from oracle import c_oracle as C, nms_ref as N
    from tdrn_b200.utils.nms_wrapper import nms
    from tdrn_b200 import ops
    # heavy suppression: everything piled on a few centres; quantised scores -> many exact ties
    rng = np.random.RandomState(11)
    c = rng.randint(0, 5, size=3000)
    xy = c[:, None] * 60.0 + rng.rand(3000, 2) * 6
    dets = np.hstack([xy, xy + 40 + rng.rand(3000, 2) * 4, np.round(rng.rand(3000, 1), 2)]).astype(np.float32)
    assert nms(dets, 0.45, force_cpu=True) == C.cpu_nms(dets, 0.45) == N.cpu_nms(dets, 0.45)
    assert nms(dets, 0.45) == nms(dets, 0.45, force_cpu=True)          # no pair sits exactly on the threshold here
    # identical boxes with identical scores: only the lowest index survives
    d = np.tile(np.array([[5, 5, 50, 50, 0.5]], np.float32), (700, 1))
    assert nms(d, 0.45) == [0]
    # IoU exactly at the threshold is suppressed (>=, cpu_nms.pyx:65): boxes 0..9 and 5..14 (+1 convention) -> 5/15
    e = np.array([[0, 0, 9, 0, 0.9], [5, 0, 14, 0, 0.8]], np.float32)
    assert nms(e, 1.0 / 3.0, force_cpu=True) == C.cpu_nms(e, 1.0 / 3.0)
    # ... and kept by the reference's GPU rule (`>`, nms_kernel.cu:71), which nms() runs with the default force_cpu=False:
    # IoU([0,0,9,9], [0,0,9,19]) == 0.5 exactly in fp32
    edge = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 130, 0.7]], np.float32)
    assert nms(edge, 0.5, force_cpu=True) == [0, 2] and nms(edge, 0.5) == [0, 1, 2]
    # pinned tie rule (documented deviation, utils/nms_wrapper.py): equal scores are visited lower index first
    tie = np.array([[0, 0, 10, 10, 0.5], [1, 1, 11, 11, 0.5], [50, 50, 60, 60, 0.5], [0, 0, 10, 10, 0.7]], np.float32)
    assert nms(tie, 0.45) == [3, 2]                                     # box 3 (highest score) suppresses 0 and 1
    assert nms(tie[:3], 0.45) == [0, 2] and nms(tie[:3], 0.45, force_cpu=True) == [0, 2]
    # golden from the reference's py_cpu_nms.py
    g = golden('small_cases')
    assert nms(g['nms_dets'], 0.45) == g['nms_keep_py_cpu_nms'].tolist()
    # device API with early exit
    dets = _rand_dets(5000, 12)
    keep, num = ops.nms_device(torch.from_numpy(dets).cuda(), 0.45, max_keep=200)
    n = int(num.item())
    assert n == 200 and keep[:n].cpu().tolist() == C.cpu_nms(dets, 0.45)[:200]


def test_decode_matches_reference_golden(golden):
    from tdrn_b200 import ops
    from tdrn_b200.layers.functions import PriorBox
    from tdrn_b200.data import mb_cfg
    from oracle.make_golden import syn_inputs
    g = golden('small_cases')
    loc, arm, conf = syn_inputs()
    pri = PriorBox(mb_cfg['VOC_320']).forward()
    dec = ops.decode(loc.cuda(), pri.cuda(), arm.cuda()).cpu().numpy()
    ref = g['syn_decode']
    # identical op order; only expf may differ from the CPU libm by an ulp or two
    assert np.abs(dec - ref).max() <= 8 * np.spacing(np.float32(np.abs(ref).max()))
    dec1 = ops.decode(loc.cuda(), pri.cuda(), None).cpu().numpy()
    from oracle import c_oracle as C
    assert np.abs(dec1 - C.decode(loc.numpy(), pri.numpy(), None)).max() <= 8 * np.spacing(np.float32(np.abs(dec1).max()))


@pytest.mark.parametrize('top_k,conf_t,nms_t,use_arm,scale', [
    (200, 0.01, 0.45, True, [320.] * 4), (50, 0.05, 0.3, False, [500., 375., 500., 375.]), (5, 0.2, 0.45, True, [320.] * 4)])
def test_detect_bit_exact_given_decoded_boxes(top_k, conf_t, nms_t, use_arm, scale):
    """Detect output == oracle Detect run on the boxes the GPU decoded (isolates the <=2 ulp expf)."""
    from oracle import c_oracle as C
    from oracle.make_golden import syn_inputs
    from tdrn_b200 import ops
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    loc, arm, conf = syn_inputs()
    if not use_arm:
        loc = loc * 0.1
    pri = PriorBox(mb_cfg['VOC_320']).forward()
    det = Detect(21, 0, top_k, conf_t, nms_t)
    out = det.forward(loc.cuda(), conf.cuda(), pri.cuda(), arm_loc_data=arm.cuda() if use_arm else None,
                      scale=torch.tensor(scale))
    assert tuple(out.shape) == (2, 21, top_k, 5) and out.dtype == torch.float32
    boxes = ops.decode(loc.cuda(), pri.cuda(), arm.cuda() if use_arm else None).cpu().numpy()
    ref = C.detect(boxes, conf.numpy(), np.asarray(scale, np.float32), 21, top_k, conf_t, nms_t)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert not out[:, 0].any()


def test_detect_matches_reference_golden(golden):
    """End to end vs the arrays the reference's own Detect produced (tolerating decode ulps)."""
    from oracle.make_golden import syn_inputs
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    g = golden('small_cases')
    loc, arm, conf = syn_inputs()
    pri = PriorBox(mb_cfg['VOC_320']).forward()
    out = Detect(21, 0, 200, 0.01, 0.45).forward(loc.cuda(), conf.cuda(), pri.cuda(), arm_loc_data=arm.cuda()).cpu().numpy()
    ref = g['syn_detect']
    assert np.array_equal(out[..., 0], ref[..., 0])                     # same detections kept, same order
    assert np.abs(out[..., 1:] - ref[..., 1:]).max() < 1e-5


def test_detect_random_init_regime_and_empty():
    """Regime R (every prior is a candidate for every class) and the no-candidate case."""
    from oracle import c_oracle as C
    from tdrn_b200 import ops
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    g = torch.Generator().manual_seed(21)
    P, Cn, B = 6375, 21, 2
    pri = PriorBox(mb_cfg['VOC_320']).forward()
    loc = torch.randn(B, P, 4, generator=g) * 0.5
    arm = torch.randn(B, P, 4, generator=g) * 0.5
    conf = torch.softmax(torch.randn(B * P, Cn, generator=g) * 0.1, 1)      # ~1/21 everywhere > 0.01
    out = Detect(Cn, 0, 200, 0.01, 0.45).forward(loc.cuda(), conf.cuda(), pri.cuda(), arm_loc_data=arm.cuda())
    boxes = ops.decode(loc.cuda(), pri.cuda(), arm.cuda()).cpu().numpy()
    ref = C.detect(boxes, conf.numpy(), np.array([320.] * 4, np.float32), Cn, 200, 0.01, 0.45)
    assert np.array_equal(out.cpu().numpy(), ref)
    out = Detect(Cn, 0, 200, 0.99, 0.45).forward(loc.cuda(), conf.cuda(), pri.cuda(), arm_loc_data=arm.cuda())
    assert not out.any()


def test_detect_coco512_shape():
    from oracle import c_oracle as C
    from tdrn_b200 import ops
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    g = torch.Generator().manual_seed(22)
    P, Cn, B = 16320, 81, 1
    pri = PriorBox(mb_cfg['VOC_512_RefineDet']).forward()
    loc = torch.randn(B, P, 4, generator=g) * 0.5
    logits = torch.randn(B * P, Cn, generator=g) * 2
    logits[:, 0] += 3
    conf = torch.softmax(logits, 1)
    out = Detect(Cn, 0, 100, 0.01, 0.45).forward(loc.cuda(), conf.cuda(), pri.cuda(), scale=torch.tensor([512.] * 4))
    boxes = ops.decode(loc.cuda(), pri.cuda(), None).cpu().numpy()
    ref = C.detect(boxes, conf.numpy(), np.array([512.] * 4, np.float32), Cn, 100, 0.01, 0.45)
    assert np.array_equal(out.cpu().numpy(), ref)


def test_collect_detections_matches_eval_loop():
    """tdrn_collect_detections == the reference's per-image / per-class masked_select + scale + hstack (evaluate.py:469-483),
    bit for bit, including empty segments, a non-prefix mask and the row cap."""
    from oracle.eval_ref import all_boxes_ref
    from tdrn_b200.utils.results import collect_detections, to_all_boxes
    g = torch.Generator().manual_seed(5)
    B, C, K = 5, 7, 40
    det = torch.zeros(B, C, K, 5)
    for b in range(B):
        for c in range(C):
            n = int(torch.randint(0, K + 1, (1,), generator=g)) if (b + c) % 3 else 0
            det[b, c, :n, 0] = torch.rand(n, generator=g).sort(descending=True)[0] + 0.01
            det[b, c, :n, 1:] = torch.rand(n, 4, generator=g)
    det[2, 3, 5, 0] = 0.0                                            # a hole: the mask is not a prefix
    det[:, 0] = torch.rand(B, K, 5, generator=g)                     # background rows must be ignored
    sizes = [(500, 375), (353, 500), (1280, 720), (320, 320), (64, 48)]
    ref = all_boxes_ref(det, sizes)
    rows, n = collect_detections(det.cuda(), sizes)
    got = to_all_boxes(rows, B, C)
    assert n == sum(len(ref[c][b]) for c in range(C) for b in range(B)) == rows.shape[0]
    for c in range(C):
        for b in range(B):
            if len(ref[c][b]) == 0:
                assert len(got[c][b]) == 0
            else:
                assert np.array_equal(got[c][b], ref[c][b]), (c, b)
    capped, n2 = collect_detections(det.cuda(), sizes, max_rows=17)
    assert n2 == n and capped.shape[0] == 17 and torch.equal(capped, rows[:17])


def _detect_vs_oracle(loc, conf, arm, Cn, top_k=200, conf_t=0.01, nms_t=0.45):
    from oracle import c_oracle as C
    from tdrn_b200 import ops
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    pri = PriorBox(mb_cfg['VOC_320']).forward()
    out = Detect(Cn, 0, top_k, conf_t, nms_t).forward(loc.cuda(), conf.cuda(), pri.cuda(), arm_loc_data=arm.cuda())
    boxes = ops.decode(loc.cuda(), pri.cuda(), arm.cuda()).cpu().numpy()
    ref = C.detect(boxes, conf.numpy(), np.array([320.] * 4, np.float32), Cn, top_k, conf_t, nms_t)
    assert np.array_equal(out.cpu().numpy(), ref)
    return out


def test_detect_trained_like_regime_small_segments():
    """Regime T: background logit +7, ~1-2 % of the (prior, class) scores pass the threshold -> every segment
    is a small one and takes the all-pairs path of nms_segment_kernel (nms_small_path)."""
    g = torch.Generator().manual_seed(31)
    P, Cn, B = 6375, 21, 4
    loc = torch.randn(B, P, 4, generator=g) * 0.5
    arm = torch.randn(B, P, 4, generator=g) * 0.5
    logits = torch.randn(B * P, Cn, generator=g)
    logits[:, 0] += 7.7                                                   # bench.py's regime T: ~1.2 % candidates
    conf = torch.softmax(logits, 1)
    n_cand = (conf.view(B, P, Cn)[:, :, 1:] > 0.01).sum(1)
    assert 0 < int(n_cand.max()) <= 256 and float(n_cand.float().mean()) > 20
    out = _detect_vs_oracle(loc, conf, arm, Cn)
    assert (out[:, 1:, 0, 0] > 0).any()
    _detect_vs_oracle(loc, conf, arm, Cn, top_k=7, conf_t=0.02, nms_t=0.3)     # early exit at top_k inside a chunk


def test_detect_segment_sizes_around_the_warp_kernel_limit():
    """Segments with exactly 0, 1, 31, 32, 33, 255, 256 (all-pairs path) and 257, 1025 (selection batches) candidates in ONE call, with
    ties in the scores (pinned rule: lower prior index first)."""
    g = torch.Generator().manual_seed(32)
    P, Cn, B = 6375, 11, 1
    loc = torch.randn(B, P, 4, generator=g) * 0.3
    arm = torch.randn(B, P, 4, generator=g) * 0.3
    conf = torch.full((B * P, Cn), 0.001)
    sizes = [0, 1, 31, 32, 33, 255, 256, 257, 1025, 200]
    for cl, n in enumerate(sizes, start=1):
        idx = torch.randperm(P, generator=g)[:n]
        sc = 0.02 + 0.9 * torch.rand(n, generator=g)
        sc = torch.round(sc * 64) / 64 if cl == 10 else sc                # class 10: heavy ties
        conf[idx, cl] = sc
    _detect_vs_oracle(loc, conf, arm, Cn)
    _detect_vs_oracle(loc, conf, arm, Cn, top_k=600)                       # kept list longer than a chunk, > 512 rows
