"""MobileNet 16-bit path: half-precision trunk against the bf16 trunk (TDRN_MOBILE_BF16=1), both against the oracle with the
oracle's offsets given -- the numbers behind tests/test_gpu_models.py::test_bf16_heads_with_reference_offsets[drn_mobilenet320].
Test infrastructure (imports oracle/)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import torch
from oracle import model_ref as M
from oracle.make_golden import CASES, make_input
import importlib

def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())

mod_name, spec_fn, build_kw, spec_kw, _ = CASES['drn_mobilenet320']
mod = importlib.import_module('tdrn_b200.model.dualrefinedet_mobilenet')
from oracle.make_golden import SEED_W
sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
x = make_input(2, 320, seed=9)
with torch.no_grad():
    ref = M.drn_mobilenet_forward(sd, x, **spec_kw)
for mode in ('half', 'bf16'):
    os.environ['TDRN_MOBILE_BF16'] = '1' if mode == 'bf16' else '0'
    net = mod.build_net('test', **build_kw)
    net.load_state_dict(sd)
    net = net.eval().cuda().set_precision('bf16')
    with torch.no_grad():
        out = net(x.cuda(), _offsets=([o.cuda() for o in ref[1]], None))
        out2 = net(x.cuda())
    torch.cuda.synchronize()
    print('%-5s trunk, oracle offsets given: arm_loc %.3e odm_loc %.3e conf %.3e | own offsets: odm_loc %.3e conf %.3e' % (
        mode, rel(out[0].cpu().numpy(), ref[0].numpy()), rel(out[2].cpu().numpy(), ref[2].numpy()), rel(out[3].cpu().numpy(), ref[3].numpy()),
        rel(out2[2].cpu().numpy(), ref[2].numpy()), rel(out2[3].cpu().numpy(), ref[3].numpy())))
