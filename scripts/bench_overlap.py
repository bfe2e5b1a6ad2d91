"""Experiment: consecutive steps on two alternating streams / graph instances, so that step i+1's trunk can fill the
SMs the latency-bound tail of step i leaves idle.  Prints ms per step for 1 and 2 in-flight steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from tdrn_b200.layers.functions import Detect, PriorBox
from tdrn_b200.data import mb_cfg
from tdrn_b200.utils.synthetic import frames

dev = torch.device('cuda')
net = Bn.build_synthetic_net().to(dev).set_precision('bf16')
det = Detect(21, 0, 200, 0.01, 0.45)
pri = PriorBox(mb_cfg['VOC_320']).forward().to(dev)
xs = [frames(32, 320, seed=100 + i).to(dev) for i in range(4)]


def hot(x):
    a, _, l, c = net(x)
    return det.forward(l, c, pri, arm_loc_data=a)


def build(n_inflight):
    streams = [torch.cuda.Stream(dev) for _ in range(n_inflight)]
    sx = [torch.empty_like(xs[0]) for _ in range(n_inflight)]
    graphs, outs = [], []
    for k in range(n_inflight):
        with torch.cuda.stream(streams[k]), torch.no_grad():
            for _ in range(2):
                hot(xs[0])
            streams[k].synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[k]):
                outs.append(hot(sx[k]))
            graphs.append(g)
    return streams, sx, graphs, outs


for n in (1, 2, 3):
    streams, sx, graphs, outs = build(n)
    torch.cuda.synchronize()

    def run(steps):
        for i in range(steps):
            k = i % n
            with torch.cuda.stream(streams[k]):
                sx[k].copy_(xs[i % 4], non_blocking=True)
                graphs[k].replay()
    run(12)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_stream(streams[0])
    steps = 30
    run(steps)
    for s in streams[1:]:
        streams[0].wait_stream(s)
    e1.record(streams[0])
    torch.cuda.synchronize()
    print('in-flight steps %d: %.4f ms per step' % (n, e0.elapsed_time(e1) / steps))
