"""Host-side logic that needs no GPU: stale packed-weight detection of the detector mirrors."""
import copy

import torch


def test_param_stamp_sees_in_place_updates_and_rebinding():
    from tdrn_b200.model import dualrefinedet_vggbn as V
    net = V.build_net('test', 320, 21)
    s0 = net._param_stamp()
    assert net._param_stamp() == s0
    with torch.no_grad():
        next(net.parameters()).mul_(1.0)                       # in-place write: version counter moves
    s1 = net._param_stamp()
    assert s1 != s0
    bn = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)][0]
    bn.running_mean.add_(0.0)                                  # buffers too (manual BN-statistics edit)
    s2 = net._param_stamp()
    assert s2 != s1
    bn.running_mean.data.add_(0.0)                             # writes through `.data` bypass the version counters:
    assert net._param_stamp() == s2                            # undetectable by design -> net.refresh() is the documented call
    net.load_state_dict(net.state_dict())                      # invalidates the cached tensor list as well
    assert net.__dict__['_stamp_tensors'] is None and net._engine is None
    clone = copy.deepcopy(net)
    assert clone._engine is None and clone._param_stamp() != net._param_stamp()     # own storage
    assert net.refresh() is net
