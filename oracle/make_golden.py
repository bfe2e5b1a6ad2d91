"""Generate tests/golden/*.npz by executing the REAL reference Python (imported in place from
/root/reference through oracle/ref_shim.py) on seeded inputs and seeded weights.

    python -m oracle.make_golden            (build container only; /root/reference must exist)

TEST INFRASTRUCTURE ONLY.  The fixtures pin
  * the functional restatements in oracle/model_ref.py and oracle/detect_ref.py (CPU tests), and
  * the CUDA path (pytest -m gpu), which regenerates the same weights/inputs from the seeds with
    oracle.model_ref.make_state_dict (the 34M-parameter state dict is too large to commit) and
    compares against the arrays stored here.
Weight generation is deterministic for a given torch version (CPU mt19937); each fixture stores a
float64 checksum of the state dict so a drifting RNG is detected instead of mis-reported as a kernel bug.
The native deformable conv inside the reference modules is the oracle restatement (the CUDA/THC
extension cannot be built), so for ConvOffset2d these fixtures pin "reference graph + restated op".
"""
import os

import numpy as np
import torch

from . import ref_shim, model_ref as M, detect_ref as D

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')

CASES = {
    # name: (builder, spec fn, kwargs for reference build_net, kwargs for spec, prior stride to store)
    'drn_vgg320_multihead': ('drn_vgg', M.param_spec_drn_vgg,
                             dict(size=320, num_classes=21, def_groups=1, bn=True, multihead=True),
                             dict(num_classes=21, def_groups=1, bn=True, multihead=True), 1),
    'drn_vgg320_single': ('drn_vgg', M.param_spec_drn_vgg,
                          dict(size=320, num_classes=21, def_groups=1, bn=True, multihead=False),
                          dict(num_classes=21, def_groups=1, bn=True, multihead=False), 7),
    'drn_mobilenet320': ('drn_mobilenet', M.param_spec_drn_mobilenet,
                         dict(size=320, num_classes=21, def_groups=1, multihead=False),
                         dict(num_classes=21, def_groups=1, multihead=False), 7),
    'refinedet_vgg320': ('refinedet_vgg', M.param_spec_refinedet_vgg,
                         dict(size=320, num_classes=21, use_refine=True, bn=False),
                         dict(num_classes=21, use_refine=True, bn=False), 7),
}
SEED_W, SEED_X = 0, 1


def make_input(b, size, seed=SEED_X):
    return torch.randn(b, 3, size, size, generator=torch.Generator().manual_seed(seed))


def run_case(ns, name):
    mod_name, spec_fn, build_kw, spec_kw, stride = CASES[name]
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    net = getattr(ns, mod_name).build_net('test', **build_kw)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = make_input(1, build_kw['size'])
    with torch.no_grad():
        out = net(x)
    arm_loc, odm_loc, conf = out[0], out[2], out[3]
    rec = dict(arm_loc=arm_loc[0, ::stride].numpy(), odm_loc=odm_loc[0, ::stride].numpy(),
               conf=conf[::stride].numpy(), stride=np.int64(stride),
               sd_checksum=np.float64(M.state_dict_checksum(sd)), x_checksum=np.float64(x.double().abs().sum()))
    if name == 'drn_vgg320_multihead':
        # Detect on the reference's own outputs (detection.py) with the canonical thresholds
        pri = ns.PriorBox(D.VOC_320).forward()
        det = ns.Detect(21, 0, 200, 0.01, 0.45).forward(odm_loc, conf, pri, arm_loc_data=arm_loc)
        rec['detect'] = det.numpy()
        rec['offset0'] = out[1][0][0].numpy()
        rec['offset3'] = out[1][3][0].numpy()
    return rec


def tdrn_case(ns):
    """TDRN key-frame step: static SSD4Scale (ret_loc) drives the temporal net's offsets (evaluate_trn.py:450-466)."""
    C = 31
    sd_s = M.make_state_dict(M.param_spec_ssd4scale_vgg(C, bn=True, deform=False), SEED_W)
    sd_t = M.make_state_dict(M.param_spec_ssd4scale_vgg(C, bn=True, deform=True), SEED_W + 1)
    static = ns.ssd4scale_vgg.build_net('test', 320, C, bn=True, deform=False)
    temporal = ns.ssd4scale_vgg.build_net('test', 320, C, bn=True, deform=True)
    static.load_state_dict(sd_s, strict=True); temporal.load_state_dict(sd_t, strict=True)
    static.eval(); temporal.eval()
    x = make_input(1, 320)
    with torch.no_grad():
        s_out = static(x, ret_loc=True)
        t_out = temporal(x, ref_loc=s_out[2], offset_list=[], ret_off=True)
    st = 7
    return dict(static_loc=s_out[0][0, ::st].numpy(), static_conf=s_out[1][::st].numpy(),
                temporal_loc=t_out[0][0, ::st].numpy(), temporal_conf=t_out[1][::st].numpy(),
                offset0=t_out[2][0][0, :, ::4, ::4].numpy(), stride=np.int64(st),
                sd_checksum=np.float64(M.state_dict_checksum(sd_s) + M.state_dict_checksum(sd_t)))


def tdrn_mobile_case(ns):
    """The same key-frame step with the MobileNet TDRN pair (model/ssd4scale_mobile.py; evaluate_trn.py:537)."""
    C = 31
    sd_s = M.make_state_dict(M.param_spec_ssd4scale_mobile(C, deform=False), SEED_W)
    sd_t = M.make_state_dict(M.param_spec_ssd4scale_mobile(C, deform=True), SEED_W + 1)
    static = ns.ssd4scale_mobile.build_net('test', 320, C, deform=False)
    temporal = ns.ssd4scale_mobile.build_net('test', 320, C, deform=True)
    static.load_state_dict(sd_s, strict=True); temporal.load_state_dict(sd_t, strict=True)
    static.eval(); temporal.eval()
    x = make_input(1, 320)
    with torch.no_grad():
        s_out = static(x, ret_loc=True)
        t_out = temporal(x, ref_loc=s_out[2], offset_list=[], ret_off=True)
    st = 7
    return dict(static_loc=s_out[0][0, ::st].numpy(), static_conf=s_out[1][::st].numpy(),
                temporal_loc=t_out[0][0, ::st].numpy(), temporal_conf=t_out[1][::st].numpy(),
                offset0=t_out[2][0][0, :, ::4, ::4].numpy(), stride=np.int64(st),
                sd_checksum=np.float64(M.state_dict_checksum(sd_s) + M.state_dict_checksum(sd_t)))


def syn_inputs(B=2, P=6375, C=21, seed=7):
    """Synthetic Detect inputs (regime T: background-dominated softmax), regenerated by the tests."""
    g = torch.Generator().manual_seed(seed)
    loc = torch.randn(B, P, 4, generator=g)
    arm = torch.randn(B, P, 4, generator=g) * 0.5
    logits = torch.randn(B * P, C, generator=g) * 2
    logits[:, 0] += 4.0
    return loc, arm, torch.softmax(logits, 1)


def small_cases(ns):
    """PriorBox tables, decode and Detect on synthetic inputs, straight from the reference's files."""
    rec = {}
    for cname, cfg in (('VOC_320', D.VOC_320), ('VOC_512_RefineDet', D.VOC_512_RefineDet)):
        rec['priors_' + cname] = ns.PriorBox(cfg).forward().numpy()
    pri = ns.PriorBox(D.VOC_320).forward()
    loc, arm, conf = syn_inputs()
    C = conf.shape[1]
    B = loc.shape[0]
    g = torch.Generator().manual_seed(8)
    rec['syn_decode'] = torch.stack([ns.box_utils.decode(loc[i], ns.box_utils.center_size(
        ns.box_utils.decode(arm[i], pri, [0.1, 0.2])), [0.1, 0.2]) for i in range(B)]).numpy()
    rec['syn_detect'] = ns.Detect(C, 0, 200, 0.01, 0.45).forward(loc, conf, pri, arm_loc_data=arm).numpy()
    rec['syn_detect_noarm'] = ns.Detect(C, 0, 50, 0.05, 0.3).forward(
        loc * 0.1, conf, pri, scale=torch.tensor([500., 375., 500., 375.])).numpy()
    # L2Norm module of the reference
    x = torch.randn(2, 16, 5, 7, generator=g)
    l2 = ns.L2Norm(16, 10)
    with torch.no_grad():
        l2.weight.copy_(torch.rand(16, generator=g) * 10)
        rec['l2_x'], rec['l2_w'], rec['l2_y'] = x.numpy(), l2.weight.numpy().copy(), l2(x).numpy()
    # reference's importable numpy NMS (py_cpu_nms.py; '>'-suppress variant)
    rng = np.random.RandomState(3)
    xy = rng.rand(400, 2) * 300
    wh = rng.rand(400, 2) * 80 + 4
    dets = np.hstack([xy, xy + wh, rng.rand(400, 1)]).astype(np.float32)
    rec['nms_dets'] = dets
    rec['nms_keep_py_cpu_nms'] = np.asarray(ns.py_cpu_nms(dets, 0.45), dtype=np.int64)
    return rec


def main(only=()):
    """``python -m oracle.make_golden [fixture ...]``: all fixtures, or only the named ones."""
    ns = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    jobs = [(name, (lambda name=name: run_case(ns, name))) for name in CASES]
    jobs += [('tdrn_vgg320_keyframe', lambda: tdrn_case(ns)), ('tdrn_mobile320_keyframe', lambda: tdrn_mobile_case(ns)),
             ('small_cases', lambda: small_cases(ns))]
    unknown = set(only) - set(n for n, _ in jobs)
    if unknown:
        raise SystemExit('unknown fixture(s): %s' % ', '.join(sorted(unknown)))
    for name, fn in jobs:
        if only and name not in only:
            continue
        rec = fn()
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
        print(name, {k: getattr(v, 'shape', v) for k, v in rec.items()})


if __name__ == '__main__':
    import sys
    main(tuple(sys.argv[1:]))
