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
    boxes = torch.stack([D.decode_two_stage(loc[i], pri, arm[i] if use_arm else None) for i in range(B)]).numpy()
    got_c = Cc.detect(boxes, conf.numpy(), np.array([320.] * 4, np.float32), C, top_k, conf_t, nms_t)
    assert np.array_equal(got_c, ref)
    assert (ref[:, 1:, :, 0] > 0).any()
