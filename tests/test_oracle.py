"""CPU tests: the oracle against the committed golden vectors (generated from the real reference by
oracle/make_golden.py), against its own C restatement, and against independent cross-checks."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_ref as M, detect_ref as D, deform_conv_ref as R, nms_ref as N, c_oracle as C
from oracle.make_golden import CASES, SEED_W, make_input, syn_inputs
from conftest import rel_err


@pytest.mark.parametrize('name', ['drn_vgg320_multihead', 'drn_vgg320_single', 'drn_mobilenet320', 'refinedet_vgg320'])
def test_model_restatement_matches_reference_golden(golden, name):
    g = golden(name)
    mod_name, spec_fn, build_kw, spec_kw, stride = CASES[name]
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    assert abs(M.state_dict_checksum(sd) - float(g['sd_checksum'])) < 1e-6 * float(g['sd_checksum']), 'weight RNG drifted'
    x = make_input(1, 320)
    assert abs(float(x.double().abs().sum()) - float(g['x_checksum'])) < 1e-6 * float(g['x_checksum'])
    fwd = {'drn_vgg': M.drn_vgg_forward, 'drn_mobilenet': M.drn_mobilenet_forward,
           'refinedet_vgg': M.refinedet_vgg_forward}[mod_name]
    out = fwd(sd, x, **spec_kw)
    for key, t in (('arm_loc', out[0][0]), ('odm_loc', out[2][0]), ('conf', out[3])):
        # same op sequence as the reference modules on the same CPU build -> tiny or zero difference
        assert rel_err(t[::stride].numpy(), g[key]) < 1e-5, key


def test_tdrn_restatement_matches_reference_golden(golden):
    g = golden('tdrn_vgg320_keyframe')
    sd_s = M.make_state_dict(M.param_spec_ssd4scale_vgg(31, bn=True, deform=False), SEED_W)
    sd_t = M.make_state_dict(M.param_spec_ssd4scale_vgg(31, bn=True, deform=True), SEED_W + 1)
    x = make_input(1, 320)
    s = M.ssd4scale_vgg_forward(sd_s, x, 31, bn=True, deform=False, ret_loc=True)
    t = M.ssd4scale_vgg_forward(sd_t, x, 31, bn=True, deform=True, ref_loc=s[2], ret_off=True)
    st = int(g['stride'])
    assert rel_err(s[0][0, ::st].numpy(), g['static_loc']) < 1e-5
    assert rel_err(s[1][::st].numpy(), g['static_conf']) < 1e-5
    assert rel_err(t[0][0, ::st].numpy(), g['temporal_loc']) < 1e-5
    assert rel_err(t[1][::st].numpy(), g['temporal_conf']) < 1e-5
    assert rel_err(t[2][0][0, :, ::4, ::4].numpy(), g['offset0']) < 1e-5


def test_tdrn_mobile_restatement_matches_reference_golden(golden):
    """MobileNet TDRN pair (model/ssd4scale_mobile.py): restatement vs the reference's own modules."""
    g = golden('tdrn_mobile320_keyframe')
    sd_s = M.make_state_dict(M.param_spec_ssd4scale_mobile(31, deform=False), SEED_W)
    sd_t = M.make_state_dict(M.param_spec_ssd4scale_mobile(31, deform=True), SEED_W + 1)
    chk = M.state_dict_checksum(sd_s) + M.state_dict_checksum(sd_t)
    assert abs(chk - float(g['sd_checksum'])) < 1e-6 * float(g['sd_checksum']), 'weight RNG drifted'
    x = make_input(1, 320)
    s = M.ssd4scale_mobile_forward(sd_s, x, 31, deform=False, ret_loc=True)
    t = M.ssd4scale_mobile_forward(sd_t, x, 31, deform=True, ref_loc=s[2], ret_off=True)
    assert tuple(s[0].shape) == (1, 6375, 4) and tuple(s[1].shape) == (6375, 31)
    assert [tuple(m.shape) for m in s[2]] == [(1, 12, 40, 40), (1, 12, 20, 20), (1, 12, 10, 10), (1, 12, 5, 5)]
    st = int(g['stride'])
    assert rel_err(s[0][0, ::st].numpy(), g['static_loc']) < 1e-5
    assert rel_err(s[1][::st].numpy(), g['static_conf']) < 1e-5
    assert rel_err(t[0][0, ::st].numpy(), g['temporal_loc']) < 1e-5
    assert rel_err(t[1][::st].numpy(), g['temporal_conf']) < 1e-5
    assert rel_err(t[2][0][0, :, ::4, ::4].numpy(), g['offset0']) < 1e-5
    # cached offsets on non-key frames (evaluate_trn.py:460-462): same result as recomputing them
    t2 = M.ssd4scale_mobile_forward(sd_t, x, 31, deform=True, offset_list=t[2])
    assert len(t2) == 2 and torch.equal(t2[0], t[0]) and torch.equal(t2[1], t[1])


def test_prior_box_bit_exact(golden):
    g = golden('small_cases')
    for name, cfg in (('VOC_320', D.VOC_320), ('VOC_512_RefineDet', D.VOC_512_RefineDet)):
        ref = g['priors_' + name]
        assert np.array_equal(D.prior_box(cfg).numpy(), ref)
        assert np.array_equal(C.prior_box(cfg), ref)
    assert D.prior_box(D.VOC_320).shape == (6375, 4) and D.prior_box(D.VOC_512_RefineDet).shape == (16320, 4)


def test_decode_and_detect_match_reference_golden(golden):
    g = golden('small_cases')
    loc, arm, conf = syn_inputs()
    pri = D.prior_box(D.VOC_320)
    dec = torch.stack([D.decode_two_stage(loc[i], pri, arm[i]) for i in range(2)])
    assert np.array_equal(dec.numpy(), g['syn_decode'])
    assert np.array_equal(D.detect(loc, conf, pri, arm).numpy(), g['syn_detect'])
    out = D.detect(loc * 0.1, conf, pri, None, torch.tensor([500., 375., 500., 375.]), 21, 50, 0.05, 0.3)
    assert np.array_equal(out.numpy(), g['syn_detect_noarm'])
    # C restatement: identical given the same decoded boxes; its own decode differs by <= 2 ulp (libm expf)
    cdet = C.detect(g['syn_decode'], conf.numpy(), np.array([320.] * 4, np.float32), 21, 200, 0.01, 0.45)
    assert np.array_equal(cdet, g['syn_detect'])
    cdec = C.decode(loc.numpy(), pri.numpy(), arm.numpy())
    assert np.abs(cdec - g['syn_decode']).max() <= 4 * np.spacing(np.abs(g['syn_decode']).max())


def test_detect_drn_golden(golden):
    g = golden('drn_vgg320_multihead')
    pri = D.prior_box(D.VOC_320)
    out = D.detect(torch.from_numpy(g['odm_loc'])[None], torch.from_numpy(g['conf']), pri,
                   torch.from_numpy(g['arm_loc'])[None])
    assert np.array_equal(out.numpy(), g['detect'])


def test_nms_variants_agree(golden):
    g = golden('small_cases')
    dets = g['nms_dets']
    k_np = N.cpu_nms(dets, 0.45)
    assert k_np == C.cpu_nms(dets, 0.45)
    # the reference's importable py_cpu_nms.py differs only at ovr == thresh exactly
    assert k_np == g['nms_keep_py_cpu_nms'].tolist()
    assert C.cpu_nms(dets, 0.45, 17) == k_np[:17]
    assert N.cpu_nms(dets[:0], 0.45) == [] and C.cpu_nms(dets[:1], 0.45) == [0]
    # ties: equal scores keep the lower index first (pinned rule)
    d = np.array([[0, 0, 10, 10, 0.5], [100, 100, 110, 110, 0.5], [0, 0, 10, 10, 0.5]], np.float32)
    assert N.cpu_nms(d, 0.5) == [0, 1] and C.cpu_nms(d, 0.5) == [0, 1]


def test_l2norm_golden(golden):
    g = golden('small_cases')
    y = M.l2norm(torch.from_numpy(g['l2_x']), torch.from_numpy(g['l2_w']))
    assert np.array_equal(y.numpy(), g['l2_y'])


@pytest.mark.parametrize('k,pad,dg,stride,dil', [(3, 1, 1, 1, 1), (3, 1, 2, 1, 1), (5, 2, 1, 1, 1), (3, 1, 4, 2, 1), (3, 2, 1, 1, 2)])
def test_deform_conv_torch_vs_c_restatement(k, pad, dg, stride, dil):
    g = torch.Generator().manual_seed(k * 10 + dg)
    x = torch.randn(2, 8, 9, 11, generator=g)
    w = torch.randn(6, 8, k, k, generator=g)
    ho, wo = R.output_size(9, 11, k, k, stride, pad, dil)
    off = torch.randn(2, dg * 2 * k * k, ho, wo, generator=g) * 3     # many samples cross the border
    a = R.deform_conv_forward(x, off, w, stride, pad, dil, dg).numpy()
    b = C.deform_conv_forward(x.numpy(), off.numpy(), w.numpy(), stride, pad, dil, dg)
    assert rel_err(a, b) < 1e-5


def test_deform_conv_zero_offset_is_conv2d():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 9, 11, generator=g)
    w = torch.randn(6, 8, 3, 3, generator=g)
    off = torch.zeros(2, 18, 9, 11)
    assert rel_err(R.deform_conv_forward(x, off, w, 1, 1, 1, 1).numpy(), F.conv2d(x, w, None, 1, 1).numpy()) < 1e-5


def test_deform_conv_border_rules():
    """Single-pixel probes of the two rules that differ from torchvision."""
    x = torch.arange(16, dtype=torch.float32).view(1, 1, 4, 4) + 1
    w = torch.ones(1, 1, 1, 1)
    def at(dy, dx):
        off = torch.zeros(1, 2, 4, 4); off[0, 0] = dy; off[0, 1] = dx
        return R.deform_conv_forward(x, off, w, 1, 0, 1, 1)[0, 0]
    assert at(-0.5, 0.0)[0, 0] == 0                 # (-1,0): zero, not interpolated with the padding
    assert at(0.5, 0.0)[3, 2] == x[0, 0, 3, 2]      # [H-1,H): replicates the last row
    assert at(0.0, 0.75)[1, 3] == x[0, 0, 1, 3]     # [W-1,W): replicates the last column
    assert at(1.0, 0.0)[3, 0] == 0                  # h == H is outside


def test_deform_conv_interior_matches_torchvision():
    tv = pytest.importorskip('torchvision.ops')
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 12, 12, generator=g)
    w = torch.randn(3, 4, 3, 3, generator=g)
    off = (torch.rand(1, 18, 12, 12, generator=g) - 0.5) * 0.9
    a = R.deform_conv_forward(x, off, w, 1, 1, 1, 1)
    b = tv.deform_conv2d(x, off, w, None, 1, 1, 1)
    assert rel_err(a[..., 2:-2, 2:-2].numpy(), b[..., 2:-2, 2:-2].numpy()) < 1e-5


def test_reference_native_library_builds_and_exports():
    """oracle/_ref: the reference's own .cu files compile unmodified (nvcc, sm_100a) and export the two doors."""
    import ctypes
    from oracle import build_ref
    path = build_ref.build()
    if path is None:
        pytest.skip('reference checkout absent and oracle/_ref not prebuilt')
    L = ctypes.CDLL(path)
    assert hasattr(L, 'ref_deformable_im2col') and hasattr(L, 'ref_gpu_nms')
