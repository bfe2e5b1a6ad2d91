"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/tdrn_b200.h declares, the host-only entry point (PriorBox) is bit-exact, the Python mirror
keeps the reference's state-dict surface and error behaviour, and nothing in the product imports oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'tdrn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tdrn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from tdrn_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), 'missing export %s' % n
    assert sorted(_lib.EXPORTS) == names
    assert L.tdrn_version() >= 100


def test_prior_box_cabi_bit_exact(golden):
    from tdrn_b200.layers.functions import PriorBox
    from tdrn_b200.data import mb_cfg
    g = golden('small_cases')
    for name in ('VOC_320', 'VOC_512_RefineDet'):
        out = PriorBox(mb_cfg[name]).forward()
        assert out.dtype == torch.float32 and np.array_equal(out.numpy(), g['priors_' + name])
    cfg = dict(mb_cfg['VOC_320'], variance=[0.1, -0.2])
    with pytest.raises(ValueError):
        PriorBox(cfg)


def test_prior_box_every_mb_cfg_entry_matches_oracle():
    """All eight dictionaries of mb_cfg (fractional min_sizes: VOC_512 / COCO_512; fractional aspect ratios, flip off:
    MOT_300) through tdrn_prior_box vs the Python restatement (itself pinned live against the reference's PriorBox)."""
    from oracle import detect_ref as D
    from tdrn_b200.layers.functions import PriorBox
    from tdrn_b200.data import mb_cfg
    want = {'VOC_300': 8732, 'VOC_300_RFB': 11620, 'VOC_320': 6375, 'VOC_512': 32756, 'MOT_300': 9700, 'COCO_300': 11620,
            'COCO_512': 32756, 'VOC_512_RefineDet': 16320}
    assert set(mb_cfg) == set(want)
    for name, cfg in mb_cfg.items():
        got = PriorBox(cfg).forward().numpy()
        assert got.shape == (want[name], 4) and np.array_equal(got, D.prior_box(cfg).numpy()), name


def test_prior_box_with_max_sizes_matches_oracle():
    from oracle import detect_ref as D
    from tdrn_b200.layers.functions import PriorBox
    cfg = {'feature_maps': [38, 19, 10, 5, 3, 1], 'min_dim': 300, 'steps': [8, 16, 32, 64, 100, 300],
           'min_sizes': [30, 60, 111, 162, 213, 264], 'max_sizes': [60, 111, 162, 213, 264, 315],
           'aspect_ratios': [[2], [2, 3], [2, 3], [2, 3], [2], [2]], 'variance': [0.1, 0.2], 'clip': True,
           'flip': True, 'name': 'VOC_300'}
    assert np.array_equal(PriorBox(cfg).forward().numpy(), D.prior_box(cfg).numpy())
    cfg2 = dict(cfg, clip=False, flip=False)
    assert np.array_equal(PriorBox(cfg2).forward().numpy(), D.prior_box(cfg2).numpy())


def test_error_codes_and_messages():
    from tdrn_b200 import _lib
    L = _lib.lib()
    num = ctypes.c_int(0)
    rc = L.tdrn_prior_box(0, 0, None, None, None, None, None, None, 1, 1, None, ctypes.byref(num))
    assert rc == -1 and b'tdrn_prior_box' in L.tdrn_last_error()
    with pytest.raises(_lib.TdrnError):
        _lib.check(rc, 'tdrn_prior_box')


def test_detect_constructor_contract():
    from tdrn_b200.layers.functions import Detect
    with pytest.raises(ValueError):
        Detect(21, 0, 200, 0.01, 0.0)           # layers/functions/detection.py:20-21
    d = Detect(21, 0, 200, 0.01, 0.45)
    assert d.variance == [0.1, 0.2] and d.top_k == 200


def test_nms_wrapper_empty_input():
    from tdrn_b200.utils.nms_wrapper import nms
    assert nms(np.zeros((0, 5), np.float32), 0.45, force_cpu=True) == []   # utils/nms_wrapper.py:26-27


def test_state_dict_surface_matches_reference_keys():
    from oracle import model_ref as M
    from tdrn_b200.model import (dualrefinedet_vggbn as V, dualrefinedet_mobilenet as MB, refinedet_vgg as R, ssd4scale_vgg as S,
                                 ssd4scale_mobile as SM)

    def chk(net, spec):
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        exp = {n: tuple(s) for n, s, _ in spec}
        assert got == exp

    chk(V.build_net('test', 320, 21, multihead=True), M.param_spec_drn_vgg(21, multihead=True))
    chk(V.build_net('test', 512, 81, bn=False, def_groups=2), M.param_spec_drn_vgg(81, bn=False, def_groups=2))
    chk(MB.build_net('test', 320, 21, multihead=True), M.param_spec_drn_mobilenet(21, multihead=True))
    chk(R.build_net('test', 320, 21, use_refine=True), M.param_spec_refinedet_vgg(21, True))
    chk(S.build_net('test', 320, 31, bn=True, deform=True), M.param_spec_ssd4scale_vgg(31, bn=True, deform=True))
    chk(S.build_net('test', 320, 31, bn=True, deform=False), M.param_spec_ssd4scale_vgg(31, bn=True, deform=False))
    chk(SM.build_net('test', 320, 31, deform=True), M.param_spec_ssd4scale_mobile(31, deform=True))
    chk(SM.build_net('test', 320, 31, deform=False), M.param_spec_ssd4scale_mobile(31, deform=False))
    assert V.build_net('test', 300) is None        # dualrefinedet_vggbn.py:218-220
    net = V.build_net('test', 320, 21)
    assert (net.size, net.num_classes, net.phase) == (320, 21, 'test')


def test_no_cpu_fallback():
    from tdrn_b200.model import dualrefinedet_vggbn as V
    from tdrn_b200.model.networks import ConvOffset2d
    net = V.build_net('test', 320, 21)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 3, 320, 320))
    with pytest.raises(NotImplementedError):       # model/networks.py:632-640
        ConvOffset2d(4, 4, 3, padding=1)(torch.zeros(1, 4, 5, 5), torch.zeros(1, 18, 5, 5))
    with pytest.raises(NotImplementedError):
        V.build_net('train', 320, 21).engine()


def test_product_never_imports_oracle():
    pat = re.compile(r'^\s*(from|import)\s+\.*oracle\b', re.M)
    for d, _, files in os.walk(os.path.join(ROOT, 'tdrn_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(d, f)).read()
                assert not pat.search(txt), f
                assert 'oracle/' not in txt or f.endswith(('.cu', '.cuh')), f


def test_level_sizes_and_prior_count():
    from tdrn_b200.model._engine import level_sizes
    assert level_sizes(320) == [40, 20, 10, 5] and level_sizes(512) == [64, 32, 16, 8]
    assert 3 * sum(s * s for s in level_sizes(320)) == 6375


def test_integration_md_python_stubs_are_valid_python_and_name_real_exports():
    """Every ```python block of INTEGRATION.md compiles, and every tdrn_* symbol it binds is exported by the library."""
    src = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r"```python\n(.*?)```", src, re.S)
    assert len(blocks) >= 5
    from tdrn_b200 import _lib
    L = _lib.lib()
    for code in blocks:
        compile(code, 'INTEGRATION.md', 'exec')
        for sym in set(re.findall(r'_L\.(tdrn_[a-z0-9_]+)', code)):
            assert hasattr(L, sym), sym
    # pointers are never passed as bare integers (ctypes would truncate them to a C int)
    assert '.ctypes.data,' not in src and '.ctypes.data)' not in src


def test_dtype_codes_match_the_header():
    """The Python mirror's format codes are the header's (TDRN_F16 = IEEE half, the MobileNet trunks' format, was added in round 2)."""
    import re
    from tdrn_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'include', 'tdrn_b200.h')).read()
    codes = {m.group(1): int(m.group(2)) for m in re.finditer(r'#define\s+(TDRN_(?:F32|BF16|BF16_SPLIT|F16))\s+(\d+)', hdr)}
    assert codes == {'TDRN_F32': _lib.F32, 'TDRN_BF16': _lib.BF16, 'TDRN_BF16_SPLIT': 2, 'TDRN_F16': _lib.F16}
    assert len(set(codes.values())) == 4
