"""Attribution reports (tests/parity_tools.py) for every bf16 end-to-end case of the GPU test-suite, printed rather than
asserted: used to set the thresholds of tests/test_gpu_models.py::check_odm_attributed.  GPU box; development aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import model_ref as M
from oracle.make_golden import SEED_W, make_input
import parity_tools as PT
from tdrn_b200.model import dualrefinedet_vggbn as V
from tdrn_b200.model._engine import level_sizes


def drn(size, C, multihead, B, seed, randomize=False):
    kw = dict(num_classes=C, def_groups=1, bn=True, multihead=multihead)
    net = V.build_net('test', size, **kw)
    if randomize:
        from tdrn_b200.utils.synthetic import randomize_
        net = randomize_(net, seed=0)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    else:
        sd = M.make_state_dict(M.param_spec_drn_vgg(**kw), SEED_W)
        net.load_state_dict(sd)
    net = net.eval().cuda().set_precision('bf16')
    x = make_input(B, size, seed=seed)
    with torch.no_grad():
        out = net(x.cuda())
        src = M._vgg_trunk(sd, x, True); odm = M._fpn(sd, src)
        loc_a = [M._c(sd, 'arm_loc.%d' % k, src[k], 1, 1) for k in range(4)]
        o1 = [M._c(sd, 'offset.%d' % k, loc_a[k]) for k in range(4)]
        o2 = [M._c(sd, 'offset2.%d' % k, loc_a[k]) for k in range(4)] if multihead else None
        l_ref, c_ref = PT.odm_heads_from_offsets(sd, odm, o1, o2, C)
        sizes = [(v, v) for v in level_sizes(size)]
        maps_g = PT.arm_maps_from_flat(out[0].cpu(), sizes)
        g1 = [M._c(sd, 'offset.%d' % k, maps_g[k]) for k in range(4)]
        g2 = [M._c(sd, 'offset2.%d' % k, maps_g[k]) for k in range(4)] if multihead else None
        l_o, c_o = PT.odm_heads_from_offsets(sd, odm, g1, g2, C)
    rows = PT.drn_flipped_rows(sd, torch.cat([PT.M_flat(m) for m in loc_a], 1).view(B, -1, 4) if False else
                               torch.cat([m.permute(0, 2, 3, 1).reshape(B, -1) for m in loc_a], 1).view(B, -1, 4),
                               out[0].cpu(), sizes, multihead).reshape(-1)
    tag = 'drn %d C%d %s B%d seed %d%s' % (size, C, 'multi' if multihead else 'single', B, seed, ' randomize_' if randomize else '')
    l_g = out[2].cpu().numpy().reshape(-1, 4); c_g = out[3].cpu().numpy()
    none = np.zeros_like(rows)
    amax = max(float(m.abs().max()) for m in loc_a)
    print(tag, '| ARM err', ['%.1e' % (float((maps_g[k] - loc_a[k]).abs().max()) / amax) for k in range(4)])
    for nm, a, r, o in (('loc ', l_g, l_ref.numpy().reshape(-1, 4), l_o.numpy().reshape(-1, 4)), ('conf', c_g, c_ref.numpy(), c_o.numpy())):
        rep = PT.split_report(a, r, rows, 2e-2)
        rep2 = PT.split_report(a, o, none, 2e-2)
        print('   %s vs oracle: %s' % (nm, rep))
        print('   %s vs oracle heads fed the GPU offsets: max %.3e  beyond %d' % (nm, rep2['max_other'], rep2['n_beyond']))
    sys.stdout.flush()


drn(320, 21, True, 1, 0)
drn(320, 21, False, 1, 0)
drn(320, 21, False, 3, 5)
drn(192, 21, True, 1, 192)
drn(448, 21, True, 1, 448)
drn(704, 21, True, 1, 704)
drn(512, 81, True, 2, 21)
drn(320, 21, True, 2, 123, randomize=True)
