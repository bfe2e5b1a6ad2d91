"""Build-container-only check (skipped where /root/reference is absent, e.g. on the GPU box): the committed golden
fixtures are exactly what the reference's OWN Python produces today.  oracle/ref_shim.py imports the reference's
model/*.py and layers/* in place; oracle/make_golden.py re-runs its cases and every array must equal the committed one
bit for bit (same torch build, same seeds).  This is what pins the oracle restatements: they are compared with these
fixtures in tests/test_oracle.py, and so is the CUDA path in the -m gpu tests.

Note: ref_shim.load() patches torch.Tensor.cuda / torch.cuda.FloatTensor process-wide (the reference hard-codes them);
that is harmless here (this container has no GPU, the GPU box has no reference checkout, so the two never meet)."""
import os

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference checkout not present (GPU box)')

HERE = os.path.dirname(os.path.abspath(__file__))


def _same(rec, name):
    ref = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    assert set(rec) == set(ref.files), name
    for k in ref.files:
        assert np.array_equal(np.asarray(rec[k]), ref[k]), (name, k)


# the quick cases (a MobileNet detector, the MobileNet TDRN pair, PriorBox / decode / Detect / L2Norm / py_cpu_nms);
# `python -m oracle.make_golden` regenerates all seven fixtures the same way (VGG cases: ~20 s each on 8 cores)
@pytest.mark.parametrize('name', ['drn_mobilenet320', 'tdrn_mobile320_keyframe', 'small_cases'])
def test_committed_fixture_is_what_the_reference_produces(name):
    from oracle import make_golden as G
    ns = ref_shim.load()
    rec = {'tdrn_mobile320_keyframe': lambda: G.tdrn_mobile_case(ns), 'small_cases': lambda: G.small_cases(ns)}.get(
        name, lambda: G.run_case(ns, name))()
    _same(rec, name)


def test_reference_state_dict_keys_are_the_spec():
    """The parameter specs the oracle and the product are built from list exactly the reference modules' state-dict keys
    and shapes (load_state_dict(strict=True) inside make_golden would also fail, this says which key)."""
    from oracle import model_ref as M
    ns = ref_shim.load()
    cases = [
        (ns.drn_vgg.build_net('test', 320, 21, multihead=True), M.param_spec_drn_vgg(21, multihead=True)),
        (ns.drn_mobilenet.build_net('test', 320, 21, multihead=True), M.param_spec_drn_mobilenet(21, multihead=True)),
        (ns.refinedet_vgg.build_net('test', 320, 21, use_refine=True), M.param_spec_refinedet_vgg(21, True)),
        (ns.ssd4scale_vgg.build_net('test', 320, 31, bn=True, deform=True), M.param_spec_ssd4scale_vgg(31, bn=True, deform=True)),
        (ns.ssd4scale_mobile.build_net('test', 320, 31, deform=True), M.param_spec_ssd4scale_mobile(31, deform=True)),
        (ns.ssd4scale_mobile.build_net('test', 320, 31, deform=False), M.param_spec_ssd4scale_mobile(31, deform=False)),
    ]
    for net, spec in cases:
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        exp = {n: tuple(s) for n, s, _ in spec}
        assert got == exp, type(net).__name__


@pytest.mark.parametrize('seed,P,C,top_k,conf_t,nms_t,use_arm,bias', [
    (11, 6375, 21, 200, 0.01, 0.45, True, 4.0),     # trained-like scores on the VOC_320 priors (the canonical thresholds)
    (12, 6375, 21, 50, 0.05, 0.30, False, 4.0),     # no ARM stage, other thresholds
    (13, 700, 6, 200, 0.01, 0.45, True, 0.0),       # every prior of every class is a candidate (random-init regime)
    (14, 300, 4, 20, 0.01, 0.45, True, 0.0),        # top_k smaller than the kept list: the early exit of the product
])
def test_detect_restatements_follow_the_reference_live(seed, P, C, top_k, conf_t, nms_t, use_arm, bias):
    """The reference's own Detect.forward (detection.py:25-70, run live) == the Python restatement == the C restatement,
    bit for bit, on inputs no fixture holds."""
    import torch
    from oracle import detect_ref as D, c_oracle as Cc
    ns = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    pri = ns.PriorBox(D.VOC_320).forward()
    if P != pri.shape[0]:
        pri = pri[torch.randperm(pri.shape[0], generator=g)[:P]].contiguous()
    B = 2
    loc = torch.randn(B, P, 4, generator=g)
    arm = torch.randn(B, P, 4, generator=g) * 0.5 if use_arm else None
    logits = torch.randn(B * P, C, generator=g) * 2
    logits[:, 0] += bias
    conf = torch.softmax(logits, 1)
    ref = ns.Detect(C, 0, top_k, conf_t, nms_t).forward(loc, conf, pri, arm_loc_data=arm).numpy()
    got = D.detect(loc, conf, pri, arm, None, C, top_k, conf_t, nms_t).numpy()
    assert np.array_equal(got, ref)
    # the evaluation drivers pass the image size as NMS scale (evaluate.py:463: scale=[w, h, w, h])
    sc = torch.tensor([500., 375., 500., 375.])
    ref_s = ns.Detect(C, 0, top_k, conf_t, nms_t).forward(loc, conf, pri, arm_loc_data=arm, scale=sc).numpy()
    assert np.array_equal(D.detect(loc, conf, pri, arm, sc, C, top_k, conf_t, nms_t).numpy(), ref_s)
    boxes = torch.stack([D.decode_two_stage(loc[i], pri, arm[i] if use_arm else None) for i in range(B)]).numpy()
    got_c = Cc.detect(boxes, conf.numpy(), np.array([320.] * 4, np.float32), C, top_k, conf_t, nms_t)
    assert np.array_equal(got_c, ref)
    assert (ref[:, 1:, :, 0] > 0).any()


def _reference_functions(path, names, ns):
    """Execute the named top-level function definitions of a reference file (which cannot be imported as a module:
    multi_eval.py parses the command line and opens datasets at import time) in the namespace `ns`."""
    import ast
    tree = ast.parse(open(path).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(f.name for f in fns) == sorted(names)
    exec(compile(ast.Module(fns, []), path, 'exec'), ns)
    return ns


@pytest.mark.filterwarnings('ignore::DeprecationWarning')         # np.row_stack inside the reference's bbox_vote
def test_multi_scale_merge_follows_the_reference_test_net_live(tmp_path):
    """The reference's own `test_net` (multi_eval.py:496-655: multi-scale PriorBoxes, base_transform + cv2.flip per pass,
    Detect per pass, per-class gather with un-flip / pixel scaling / size rules, bbox_vote) executed from its source at base
    size 512 (scales 320, 512, 640, 1216 x {plain, flipped}) with real cv2, the reference's PriorBox / Detect / Cython NMS and
    a stand-in network, against oracle/multi_scale_ref.multi_scale_merge fed with the same Detect outputs: every voted box
    of every class of every image, bit for bit.  (Base size 320 cannot be pinned this way: the reference's size-rule table
    names '320_706' where its scale list says 704, and its loop then pairs boxes with a stale index array.)"""
    import ast
    import os
    import pickle
    import types
    import torch
    cv2 = pytest.importorskip('cv2')
    from oracle import multi_scale_ref as MS, make_golden_preprocess as GP
    ns_ref = ref_shim.load()
    REF = os.path.join(ref_shim.REFERENCE_ROOT, 'multi_eval.py')
    cfg = {}
    for node in ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, 'data', 'config.py')).read()).body:
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict):
            try:
                exec(compile(ast.Module([node], []), 'config.py', 'exec'), cfg)
            except NameError:
                pass
    C, TOP_K, N_IMG = 3, 25, 12                      # test_net's FPS line divides by the time of images 11.. : needs > 11 images
    rng = np.random.RandomState(5)
    images = [np.clip(rng.randn(60 + 7 * i, 90 + 5 * i, 3) * 40 + 120, 0, 255).astype(np.uint8) for i in range(N_IMG)]

    class Dataset(object):
        def __len__(self):
            return N_IMG

        def pull_image(self, i):
            return images[i]

    class Timer(object):
        def tic(self):
            self.t = 0.0

        def toc(self, average=True):
            return 1.0

    calls = {'n': 0}

    def net(x):                                      # stand-in network: seeded outputs of the right shape for this scale
        s = x.shape[-1]
        P = 3 * sum(((s // st) + (1 if s % st else 0)) ** 2 for st in (8, 16, 32, 64))
        g = torch.Generator().manual_seed(1000 + calls['n'])
        calls['n'] += 1
        loc = torch.randn(1, P, 4, generator=g) * 0.5
        arm = torch.randn(1, P, 4, generator=g) * 0.3
        logits = torch.randn(P, C, generator=g) * 2
        logits[:, 0] += 8.5
        return arm, None, loc, torch.softmax(logits, 1)

    recorded = []
    ref_detect = ns_ref.Detect(C, 0, TOP_K, 0.01, 0.45)

    class Detector(object):
        def forward(self, loc, conf, priors, arm_loc_data=None):
            assert priors.shape[0] == loc.shape[1], (priors.shape, loc.shape)
            out = ref_detect.forward(loc, conf, priors, arm_loc_data=arm_loc_data)
            recorded.append(out.clone().numpy())
            return out

    captured = {}

    def get_output_dir(name, phase):
        d = os.path.join(str(tmp_path), str(abs(hash((name, phase)))))
        os.makedirs(d, exist_ok=True)
        return d

    ns = {'np': np, 'torch': torch, 'os': os, 'cv2': cv2, 'pickle': pickle, 'labelmap': ('a', 'b'), 'Timer': Timer,
          'get_output_dir': get_output_dir, 'pkl_dir': str(tmp_path), 'device': torch.device('cpu'),
          'args': types.SimpleNamespace(iteration='0', dataset_name='x', set_file_name='y', backbone='RefineDet_VGG', refine=True),
          'multi_scale': {'320': [192, 320, 384, 448, 512, 576, 704], '512': [320, 512, 640, 1216]},
          'multi_cfg': cfg['multi_cfg_512'], 'ssd_dim': 512, 'PriorBox': ns_ref.PriorBox,
          'base_transform': GP.reference_base_transform(), 'dataset_mean': (104, 117, 123),
          'evaluate_detections': lambda all_boxes, output_dir, dataset, FPS=None: captured.update(all_boxes=all_boxes),
          'print': lambda *a, **k: None}
    _reference_functions(REF, ['test_net', 'bbox_vote'], ns)
    ns['test_net'](str(tmp_path), net, Dataset(), None, TOP_K, Detector(), None)
    all_boxes = captured['all_boxes']
    assert len(recorded) == N_IMG * 8
    n_rows = 0
    for i in range(N_IMG):
        h, w, _ = images[i].shape
        passes = [(scale, flip, recorded[i * 8 + k * 2 + flip][0]) for k, scale in enumerate([320, 512, 640, 1216]) for flip in (0, 1)]
        mine = MS.multi_scale_merge(passes, C, w, h, 512)
        for j in range(1, C):
            ref = all_boxes[j][i]
            if isinstance(ref, list) and not ref:
                assert mine[j].shape[0] == 0
                continue
            assert np.asarray(ref).dtype == np.float32 or np.asarray(ref).dtype == np.float64
            assert mine[j].shape == np.asarray(ref).shape, (i, j)
            assert np.array_equal(mine[j].astype(np.float64), np.asarray(ref, dtype=np.float64)), (i, j)
            n_rows += mine[j].shape[0]
    assert n_rows > 30                               # the comparison is not vacuous


def test_tdrn_stream_and_result_scatter_follow_the_reference_test_net_live(tmp_path):
    """The reference's own video evaluation loop (evaluate_trn.py `test_net`, :416-516) executed from its source -- key-frame
    scheduling per video and per `interval`, the in-place `loose` scaling of the static regression, the offset cache
    (`ref_loc` consumed once, `offset_list` re-used until the next key frame), Detect with the static net's regression as ARM
    stage, and the per-class result scatter -- against the PRODUCT's host logic tdrn_b200.utils.tdrn_stream.TDRNStream (pure
    Python: it runs here on CPU tensors with stand-in networks) and oracle.eval_ref.all_boxes_ref: the same calls with the
    same arguments in the same order, the same detections, the same all_boxes, bit for bit."""
    import os
    import pickle
    import types
    import torch
    from oracle import eval_ref
    from tdrn_b200.utils.tdrn_stream import TDRNStream
    ns_ref = ref_shim.load()
    REF = os.path.join(ref_shim.REFERENCE_ROOT, 'evaluate_trn.py')
    C, TOP_K, P, INTERVAL, LOOSE = 4, 30, 300, 4, 0.5
    videos = ['vidA'] * 9 + ['vidB'] * 6                       # 15 frames (> 11: the FPS lines), key frames 0,4,8 | 9,13
    sizes = [(320 + 3 * i, 240 + 2 * i) for i in range(len(videos))]           # (w, h)
    pri = ns_ref.PriorBox(dict(feature_maps=[10], min_dim=320, steps=[32], min_sizes=[64], max_sizes=[], aspect_ratios=[[2]],
                               variance=[0.1, 0.2], clip=True, flip=True, name='test')).forward()
    assert pri.shape[0] == P

    def frame_tensor(i):
        return torch.full((3, 8, 8), float(i))

    def seeded(i, salt, *shape):
        return torch.randn(*shape, generator=torch.Generator().manual_seed(10000 * salt + i))

    def make_fakes(log):
        def static_net(x, ret_loc=False):
            i = int(x[0, 0, 0, 0])
            log.append(('static', i, bool(ret_loc)))
            out = [seeded(i, 1, 1, P, 4) * 0.4, torch.softmax(seeded(i, 2, P, C), 1)]
            if ret_loc:
                out.append([seeded(i, 3, 1, 12, 10, 10), seeded(i, 4, 1, 12, 5, 5)])
            return tuple(out)

        def net(x, ref_loc=list(), offset_list=list(), ret_loc=False, ret_off=False):
            i = int(x[0, 0, 0, 0])
            log.append(('net', i, [float(t.sum()) for t in ref_loc], [float(t.sum()) for t in offset_list], bool(ret_off)))
            logits = seeded(i, 5, P, C) * 2
            logits[:, 0] += 3.0
            out = [seeded(i, 6, 1, P, 4) * 0.5, torch.softmax(logits, 1)]
            if ret_off:
                out.append([seeded(i, 7, 1, 144, 10, 10), seeded(i, 8, 1, 144, 5, 5)])
            return tuple(out)
        return static_net, net

    class Dataset(object):
        def __len__(self):
            return len(videos)

        def pull_transformed_image(self, i):
            return frame_tensor(i), sizes[i][1], sizes[i][0]     # im, h, w

        def pull_img_id(self, i):
            return ('root', '%s/%06d' % (videos[i], i))

    class Timer(object):
        def tic(self):
            pass

        def toc(self, average=True):
            return 1.0

    class RecordingDetect(object):
        def __init__(self, log):
            self.det, self.log, self.outs = ns_ref.Detect(C, 0, TOP_K, 0.01, 0.45), log, []

        def forward(self, loc, conf, priors, arm_loc_data=None):
            self.log.append(('detect', float(loc.sum()), float(arm_loc_data.sum())))
            out = self.det.forward(loc, conf, priors, arm_loc_data=arm_loc_data)
            self.outs.append(out.clone())
            return out

    # --- the reference loop, from its own source
    ref_log, captured = [], {}
    s_net, t_net = make_fakes(ref_log)
    ref_detect = RecordingDetect(ref_log)

    def get_output_dir(name, phase):
        d = os.path.join(str(tmp_path), 'out')
        os.makedirs(d, exist_ok=True)
        return d

    import numpy
    ns = {'np': numpy, 'torch': torch, 'os': os, 'pickle': pickle, 'labelmap': ('a', 'b', 'c'), 'Timer': Timer,
          'get_output_dir': get_output_dir, 'pkl_dir': str(tmp_path), 'device': torch.device('cpu'), 'iteration': '0',
          'args': types.SimpleNamespace(dataset_name='x', set_file_name='y', interval=INTERVAL, display=False, deform=True,
                                        loose=LOOSE),
          'evaluate_detections': lambda all_boxes, output_dir, dataset, FPS=None: captured.update(all_boxes=all_boxes),
          'print': lambda *a, **k: None}
    _reference_functions(REF, ['test_net'], ns)
    ns['test_net'](str(tmp_path), t_net, Dataset(), ref_detect, pri, static_net=s_net)

    # --- the product's host logic on the same stand-ins
    my_log = []
    s_net2, t_net2 = make_fakes(my_log)
    my_detect = RecordingDetect(my_log)
    stream = TDRNStream(s_net2, t_net2, my_detect, pri, interval=INTERVAL, loose=LOOSE, deform=True)
    ds = Dataset()
    mine = []
    for i in range(len(videos)):
        im, h, w = ds.pull_transformed_image(i)
        mine.append(stream.step(im.unsqueeze(0), ds.pull_img_id(i)[1].split('/')[0]).clone())

    assert my_log == ref_log
    assert [e[1] for e in ref_log if e[0] == 'static'] == [0, 4, 8, 9, 13]          # key frames: per video, every `interval`
    assert len(ref_detect.outs) == len(mine) == len(videos)
    for a, b in zip(ref_detect.outs, mine):
        assert torch.equal(a, b)
    # result scatter (the same lines in evaluate.py:469-483 and evaluate_coco.py:140-159)
    ref_boxes = captured['all_boxes']
    mine_boxes = eval_ref.all_boxes_ref(torch.cat(mine, 0), sizes)
    n = 0
    for j in range(1, C):
        for i in range(len(videos)):
            r, m = ref_boxes[j][i], mine_boxes[j][i]
            if isinstance(r, list) and not r:
                assert isinstance(m, list) and not m
                continue
            assert np.asarray(r).dtype == np.float32 and np.array_equal(np.asarray(r), np.asarray(m)), (j, i)
            n += len(r)
    assert n > 100


def test_prior_box_product_and_oracles_equal_the_reference_for_every_config():
    """Every prior-box dictionary of the reference's data/config.py (SSD-300/512 with max_sizes and fractional min_sizes,
    MOT_300 with fractional aspect ratios and flip off, the RefineDet and multi-scale ones) through the reference's own
    PriorBox.forward vs the product's host entry point tdrn_prior_box and both oracle restatements: bit-exact."""
    import ast
    import os
    from oracle import detect_ref as D, c_oracle as Cc
    from tdrn_b200.layers.functions import PriorBox
    ns = ref_shim.load()
    cfg = {}
    for node in ast.parse(open(os.path.join(ref_shim.REFERENCE_ROOT, 'data', 'config.py')).read()).body:
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict):
            try:
                exec(compile(ast.Module([node], []), 'config.py', 'exec'), cfg)
            except NameError:
                pass
    names = [k for k, c in cfg.items() if isinstance(c, dict) and 'feature_maps' in c]
    assert len(names) >= 17
    for name in names:
        ref = ns.PriorBox(cfg[name]).forward().numpy()
        assert np.array_equal(PriorBox(cfg[name]).forward().numpy(), ref), name
        assert np.array_equal(D.prior_box(cfg[name]).numpy(), ref), name
        assert np.array_equal(Cc.prior_box(cfg[name]), ref), name


def test_public_signatures_equal_the_reference():
    """Parameter names, order, kinds and defaults of the drop-in surface against the reference's own definitions
    (inspect.signature on both): build_net / __init__ / forward of the five detector modules, Detect, PriorBox, nms,
    ConvOffset2d, conv_dw, vgg.  Allowed differences: the `_offsets` test hook appended to the two DualRefineDet forwards, and
    Detect.forward's `scale` default (None here, resolved to the reference's [320, 320, 320, 320] at call time)."""
    import importlib
    import inspect
    ns = ref_shim.load()
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.model import networks as N
    from tdrn_b200.utils.nms_wrapper import nms
    pairs = []
    for refmod, mine in (('drn_vgg', 'dualrefinedet_vggbn'), ('drn_mobilenet', 'dualrefinedet_mobilenet'),
                         ('refinedet_vgg', 'refinedet_vgg'), ('ssd4scale_vgg', 'ssd4scale_vgg'), ('ssd4scale_mobile', 'ssd4scale_mobile')):
        r, m = getattr(ns, refmod), importlib.import_module('tdrn_b200.model.' + mine)
        pairs.append((mine + '.build_net', r.build_net, m.build_net))
        for c in [c for c in vars(r).values() if inspect.isclass(c) and c.__module__ == r.__name__]:
            cm = getattr(m, c.__name__)                              # same class name
            pairs.append((mine + '.' + c.__name__ + '.__init__', c.__init__, cm.__init__))
            pairs.append((mine + '.' + c.__name__ + '.forward', c.forward, cm.forward))
    pairs += [('Detect.__init__', ns.Detect.__init__, Detect.__init__), ('Detect.forward', ns.Detect.forward, Detect.forward),
              ('PriorBox.__init__', ns.PriorBox.__init__, PriorBox.__init__), ('PriorBox.forward', ns.PriorBox.forward, PriorBox.forward),
              ('nms', ns.nms_wrapper.nms, nms), ('ConvOffset2d.__init__', ns.networks.ConvOffset2d.__init__, N.ConvOffset2d.__init__),
              ('ConvOffset2d.forward', ns.networks.ConvOffset2d.forward, N.ConvOffset2d.forward),
              ('conv_dw', ns.networks.conv_dw, N.conv_dw), ('vgg', ns.networks.vgg, N.vgg)]
    assert len(pairs) >= 24
    for name, a, b in pairs:
        pa = list(inspect.signature(a).parameters.values())
        pb = list(inspect.signature(b).parameters.values())
        while name.endswith('.forward') and pb and pb[-1].name.startswith('_') and pb[-1].default is not inspect.Parameter.empty:
            pb = pb[:-1]                     # private trailing keyword extensions (`_offsets`, `_sources`; default = the reference's behaviour)
        assert [(p.name, p.kind) for p in pa] == [(p.name, p.kind) for p in pb], name
        for x, y in zip(pa, pb):
            if name == 'Detect.forward' and x.name == 'scale':
                assert y.default is None and [float(v) for v in x.default] == [320.0] * 4
                continue
            assert repr(x.default) == repr(y.default), (name, x.name)


def test_state_dict_surface_equals_the_reference_over_constructor_arguments():
    """Parameter / buffer names and shapes of every detector module against the reference's own modules over a sweep of
    constructor arguments (size, classes, c7_channel, def_groups, bn, multihead, use_refine, deform).  Names and shapes must be
    identical; the ORDER of the FPN containers differs (the reference registers trans_layers / up_layers / latent_layers after
    the heads, here they follow last_layer_trans) -- irrelevant to load_state_dict, which matches by name."""
    import contextlib
    import importlib
    import io
    ns = ref_shim.load()
    combos = {
        ('drn_vgg', 'dualrefinedet_vggbn'): [dict(size=s, num_classes=c, c7_channel=c7, def_groups=dg, bn=bn, multihead=mh)
                                             for (s, c, c7, dg, bn, mh) in [(320, 21, 1024, 1, True, True), (512, 81, 1024, 1, True, False),
                                                                            (320, 21, 512, 2, False, True), (512, 31, 1024, 4, False, False)]],
        ('drn_mobilenet', 'dualrefinedet_mobilenet'): [dict(size=s, num_classes=c, def_groups=dg, multihead=mh)
                                                       for (s, c, dg, mh) in [(320, 21, 1, False), (512, 31, 4, True)]],
        ('refinedet_vgg', 'refinedet_vgg'): [dict(size=320, num_classes=c, use_refine=ur, c7_channel=c7, bn=bn, multihead=mh)
                                             for (c, ur, c7, bn, mh) in [(21, True, 1024, False, False), (81, False, 1024, True, True),
                                                                         (21, True, 512, True, True)]],
        ('ssd4scale_vgg', 'ssd4scale_vgg'): [dict(size=320, num_classes=c, c7_channel=c7, bn=bn, deform=d)
                                             for (c, c7, bn, d) in [(31, 1024, True, True), (31, 1024, True, False), (21, 512, False, True)]],
        ('ssd4scale_mobile', 'ssd4scale_mobile'): [dict(size=320, num_classes=31, c7_channel=1024, deform=d) for d in (False, True)],
    }
    n = 0
    for (refmod, mine), lst in combos.items():
        r, m = getattr(ns, refmod), importlib.import_module('tdrn_b200.model.' + mine)
        for kw in lst:
            with contextlib.redirect_stdout(io.StringIO()):
                a, b = r.build_net('test', **kw), m.build_net('test', **kw)
            sa = {k: tuple(v.shape) for k, v in a.state_dict().items()}
            sb = {k: tuple(v.shape) for k, v in b.state_dict().items()}
            assert sa == sb, (mine, kw)
            moved = {'trans_layers', 'up_layers', 'latent_layers'}
            assert [k for k in sa if k.split('.')[0] not in moved] == [k for k in sb if k.split('.')[0] not in moved], (mine, kw)
            n += 1
    assert n == 14


@pytest.mark.parametrize('mod,kw,size', [
    ('drn_vgg', dict(num_classes=21, c7_channel=512, def_groups=2, bn=False, multihead=True), 320),
    ('drn_vgg', dict(num_classes=31, c7_channel=1024, def_groups=4, bn=True, multihead=False), 192),
    ('drn_mobilenet', dict(num_classes=31, def_groups=4, multihead=True), 320),
    ('refinedet_vgg', dict(num_classes=21, use_refine=False, c7_channel=1024, bn=True, multihead=True), 320),
    ('refinedet_vgg', dict(num_classes=81, use_refine=True, c7_channel=512, bn=False, multihead=False), 448),
])
def test_model_restatements_follow_the_reference_modules_beyond_the_fixtures(mod, kw, size):
    """The functional restatements in oracle/model_ref.py (what the -m gpu tests compare the kernels with at other sizes,
    group counts and head configurations) against the reference's own modules run live: constructor arguments and input
    sizes no committed fixture covers (deformable groups 2 / 4, c7_channel 512, no BN, COCO-81, 192 / 448-pixel inputs)."""
    import contextlib
    import io
    import torch
    from oracle import model_ref as M
    from oracle.make_golden import make_input
    ns = ref_shim.load()
    spec_fn, fwd = {'drn_vgg': (M.param_spec_drn_vgg, M.drn_vgg_forward), 'drn_mobilenet': (M.param_spec_drn_mobilenet, M.drn_mobilenet_forward),
                    'refinedet_vgg': (M.param_spec_refinedet_vgg, M.refinedet_vgg_forward)}[mod]
    sd = M.make_state_dict(spec_fn(**kw), 3)
    with contextlib.redirect_stdout(io.StringIO()):
        net = getattr(ns, mod).build_net('test', 320, **kw)           # the modules are fully convolutional: any input size runs
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = make_input(1, size, seed=4)
    with torch.no_grad():
        ref = net(x)
    out = fwd(sd, x, **kw)
    assert len(out) == len(ref)
    for k, (a, b) in enumerate(zip(out, ref)):
        if b is None:                       # slot 1 of the MobileNet / RefineDet outputs (the restatement keeps the offsets there)
            continue
        if isinstance(b, (list, tuple)):
            assert len(a) == len(b) and all(torch.equal(p, q) for p, q in zip(a, b)), k
        else:
            assert torch.equal(a, b), k
