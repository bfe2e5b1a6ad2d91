"""16-bit paths with the oracle's offsets given (the numbers behind tests/test_gpu_models.py::test_bf16_heads_with_reference_offsets):
MobileNet variant with the IEEE-half trunk against the bf16 trunk (TDRN_MOBILE_BF16=1), and both detectors with the per-tap
projections of the deformable heads stored as half (TDRN_PROJ_F16=1) against bf16.  Test infrastructure (imports oracle/)."""
import os, sys, importlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
from oracle import model_ref as M
from oracle.make_golden import CASES, SEED_W, make_input


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


for case, modname, fwd in (('drn_mobilenet320', 'dualrefinedet_mobilenet', M.drn_mobilenet_forward),
                           ('drn_vgg320_multihead', 'dualrefinedet_vggbn', M.drn_vgg_forward)):
    mod_name, spec_fn, build_kw, spec_kw, _ = CASES[case]
    mod = importlib.import_module('tdrn_b200.model.' + modname)
    sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
    x = make_input(2, 320, seed=9)
    with torch.no_grad():
        ref = fwd(sd, x, **spec_kw)
        offs2 = None
        if spec_kw.get('multihead'):
            src = M._vgg_trunk(sd, x, True)
            offs2 = [M._c(sd, 'offset2.%d' % k, M._c(sd, 'arm_loc.%d' % k, src[k], 1, 1)) for k in range(4)]
    for trunk in (('half', 'bf16') if 'mobilenet' in case else ('bf16',)):
        for proj in ('half', 'bf16'):
            os.environ['TDRN_MOBILE_BF16'] = '1' if trunk == 'bf16' else '0'
            os.environ['TDRN_PROJ_F16'] = '1' if proj == 'half' else '0'
            net = mod.build_net('test', **build_kw)
            net.load_state_dict(sd)
            net = net.eval().cuda().set_precision('bf16')
            with torch.no_grad():
                out = net(x.cuda(), _offsets=([o.cuda() for o in ref[1]], [o.cuda() for o in offs2] if offs2 else None))
            torch.cuda.synchronize()
            print('%-22s trunk %-4s projections %-4s | oracle offsets given: arm_loc %.3e odm_loc %.3e conf %.3e' % (
                case, trunk, proj, rel(out[0].cpu().numpy(), ref[0].numpy()), rel(out[2].cpu().numpy(), ref[2].numpy()),
                rel(out[3].cpu().numpy(), ref[3].numpy())))
