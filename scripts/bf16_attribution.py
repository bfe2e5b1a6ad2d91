"""Where does the bf16 path's end-to-end difference on the deformable ODM heads come from?  (GPU box; development aid
behind tests/test_gpu_parity_attribution.py.)  Prints, for DualRefineDet-VGGBN-320 multihead: the ARM regression error,
the rows beyond 2e-2 with / without a tap that changed side of the map edge (tests/parity_tools.py), and the same
outputs against the oracle evaluated WITH THE GPU'S OWN offsets."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import model_ref as M
from oracle.make_golden import CASES, make_input, SEED_W
import parity_tools as PT
from tdrn_b200.model import dualrefinedet_vggbn as V

name = sys.argv[1] if len(sys.argv) > 1 else 'drn_vgg320_multihead'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mod_name, spec_fn, build_kw, spec_kw, _ = CASES[name]
sd = M.make_state_dict(spec_fn(**spec_kw), SEED_W)
net = V.build_net('test', **build_kw); net.load_state_dict(sd); net = net.eval().cuda(); net.set_precision('bf16')
x = make_input(B, 320, seed=9)
multi = spec_kw.get('multihead', False)
with torch.no_grad():
    out = net(x.cuda())
    src = M._vgg_trunk(sd, x, True); odm = M._fpn(sd, src)
    loc_a = [M._c(sd, 'arm_loc.%d' % k, src[k], 1, 1) for k in range(4)]
    o1 = [M._c(sd, 'offset.%d' % k, loc_a[k]) for k in range(4)]
    o2 = [M._c(sd, 'offset2.%d' % k, loc_a[k]) for k in range(4)] if multi else None
    l_ref, c_ref = PT.odm_heads_from_offsets(sd, odm, o1, o2, 21)
    arm_g = out[0].cpu()
    maps_g = PT.arm_maps_from_flat(arm_g, [(m.shape[2], m.shape[3]) for m in loc_a])
    amax = max(float(m.abs().max()) for m in loc_a)
    print('ARM regression max-norm error per level', [float((maps_g[k] - loc_a[k]).abs().max()) / amax for k in range(4)])
    g1 = [M._c(sd, 'offset.%d' % k, maps_g[k]) for k in range(4)]
    g2 = [M._c(sd, 'offset2.%d' % k, maps_g[k]) for k in range(4)] if multi else None
    print('returned 3x3 offsets vs recomputed from the GPU regression', [float((out[1][k].cpu() - g1[k]).abs().max()) for k in range(4)])
    fl = [PT.flipped_pixels(o1[k], g1[k], 3, 1, 1) | (PT.flipped_pixels(o2[k], g2[k], 5, 2, 1) if multi else False) for k in range(4)]
    rows = PT.flipped_rows(fl).reshape(-1)
    l_g = out[2].cpu().numpy().reshape(-1, 4); c_g = out[3].cpu().numpy()
    print('vs oracle: loc ', PT.split_report(l_g, l_ref.numpy().reshape(-1, 4), rows, 2e-2))
    print('vs oracle: conf', PT.split_report(c_g, c_ref.numpy(), rows, 2e-2))
    e = PT.row_errors(l_g, l_ref.numpy().reshape(-1, 4)); print('largest non-flipped loc rows', np.sort(e[~rows])[-8:])
    e = PT.row_errors(c_g, c_ref.numpy()); print('largest non-flipped conf rows', np.sort(e[~rows])[-8:])
    l_o, c_o = PT.odm_heads_from_offsets(sd, odm, g1, g2, 21)
    none = np.zeros_like(rows)
    print('vs oracle heads fed the GPU offsets: loc ', PT.split_report(l_g, l_o.numpy().reshape(-1, 4), none, 2e-2))
    print('vs oracle heads fed the GPU offsets: conf', PT.split_report(c_g, c_o.numpy(), none, 2e-2))
    print('oracle own response to the GPU offsets: loc ', PT.split_report(l_o.numpy().reshape(-1, 4), l_ref.numpy().reshape(-1, 4), rows, 2e-2))
    print('oracle own response to the GPU offsets: conf', PT.split_report(c_o.numpy(), c_ref.numpy(), rows, 2e-2))
