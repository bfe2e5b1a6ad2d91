"""Throughput / latency of the other BASELINE.json configurations (configs[2..4]) through the public API, for the
record in DESIGN.md -- bench.py's JSON line stays on configs[1].  CUDA-graph replay, CUDA events, synthetic inputs
resident in HBM, 20 timed replays after 5 warm-ups.

    python scripts/bench_configs.py [coco512] [mobilenet] [tdrn] [tdrn_mobile]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200.data import mb_cfg
from tdrn_b200.layers.functions import Detect, PriorBox
from tdrn_b200.utils.synthetic import frames, randomize_

dev = torch.device('cuda')


def graph_time(fn, x, iters=20, warmup=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st), torch.no_grad():
        for _ in range(3):
            fn(x)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            out = fn(x)
        for _ in range(warmup):
            g.replay()
        st.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st)
            st.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[int(len(ts) * 0.99)], out


def coco512():
    from tdrn_b200.model import dualrefinedet_vggbn as V
    B, C = 16, 81
    net = randomize_(V.build_net('test', 512, num_classes=C, def_groups=1, bn=True, multihead=True), 0).eval().to(dev)
    pri = PriorBox(mb_cfg['VOC_512_RefineDet']).forward().to(dev)
    det = Detect(C, 0, 100, 0.01, 0.45)
    x = frames(B, 512, 3).to(dev)

    def f(x):
        a, _, l, c = net(x)
        return det.forward(l, c, pri, arm_loc_data=a, scale=[512.] * 4)
    ms, p50, p99, _ = graph_time(f, x)
    return {'config': 'DualRefineDet-VGGBN 512x512 COCO-81 batch 16 per GPU (multihead), net+Detect(top_k 100)', 'ms_per_step': ms,
            'frames_per_s_per_gpu': B / ms * 1e3, 'tflops': 215.36 * B / ms}


def mobilenet():
    from tdrn_b200.model import dualrefinedet_mobilenet as M
    net = randomize_(M.build_net('test', 320, num_classes=21, def_groups=1, multihead=False), 0).eval().to(dev)
    pri = PriorBox(mb_cfg['VOC_320']).forward().to(dev)
    det = Detect(21, 0, 200, 0.01, 0.45)

    def f(x):
        a, _, l, c = net(x)
        return det.forward(l, c, pri, arm_loc_data=a)
    res = {'config': 'DualRefineDet-MobileNet 320x320 VOC-21, net+Detect'}
    ms, p50, p99, _ = graph_time(f, frames(1, 320, 5).to(dev), iters=1000)
    res['b1_latency_ms'] = {'mean': ms, 'p50': p50, 'p99': p99}
    ms, _, _, _ = graph_time(f, frames(64, 320, 6).to(dev))
    res['b64'] = {'ms_per_step': ms, 'frames_per_s': 64 / ms * 1e3}
    return res


def _tdrn(S, kw, label, gflop_per_clip):
    """16-frame clips, key-frame interval 4: static net on frames 0,4,8,12 of every clip (ret_loc), temporal net (dg = 8
    deformable heads, offsets from the key frame's regression) on all frames, Detect with the key frame's arm_loc.  One clip
    per step (the latency configuration) and two clips per step (B = 32 temporal forwards: the throughput configuration)."""
    C, T, K = 31, 16, 4
    stat = randomize_(S.build_net('test', 320, num_classes=C, deform=False, **kw), 0).eval().to(dev)
    temp = randomize_(S.build_net('test', 320, num_classes=C, deform=True, **kw), 1).eval().to(dev)
    pri = PriorBox(mb_cfg['VOC_320']).forward().to(dev)
    det = Detect(C, 0, 200, 0.01, 0.45)

    def f(x):
        keys = x[::K]                                             # key frames (clips are stacked along the batch)
        s_loc, s_conf, loc_maps = stat(keys, ret_loc=True)
        ref = [m.repeat_interleave(K, 0) for m in loc_maps]       # every frame uses its key frame's regression
        out = temp(x, ref_loc=ref, ret_off=True)
        arm = s_loc.repeat_interleave(K, 0)
        return det.forward(out[0], out[1], pri, arm_loc_data=arm)
    res = {'config': 'TDRN %s-320 VID-31, 16-frame clips, key-frame interval 4 (4 static + 16 temporal forwards + Detect per clip)' % label}
    for clips in (1, 2):
        ms, _, _, _ = graph_time(f, frames(T * clips, 320, 7).to(dev))
        res['clips_per_step_%d' % clips] = {'ms_per_step': ms, 'frames_per_s_per_gpu': T * clips / ms * 1e3}
        if gflop_per_clip:
            res['clips_per_step_%d' % clips]['tflops'] = gflop_per_clip * clips / ms
    return res


def tdrn():
    from tdrn_b200.model import ssd4scale_vgg as S
    return _tdrn(S, dict(bn=True), 'VGGBN', 1321.0)


def tdrn_mobile():
    """The MobileNet TDRN pair (model/ssd4scale_mobile.py; `evaluate_trn.py:537`)."""
    from tdrn_b200.model import ssd4scale_mobile as S
    return _tdrn(S, {}, 'MobileNet', None)


if __name__ == '__main__':
    which = sys.argv[1:] or ['coco512', 'mobilenet', 'tdrn', 'tdrn_mobile']
    for w in which:
        try:
            print(json.dumps({w: globals()[w]()}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({w: {'error': repr(e)[:300]}}), flush=True)
