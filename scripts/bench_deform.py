"""Micro-benchmark of the deformable head at the four pyramid levels of DualRefineDet-VGGBN-320 b32 (multihead
3x3 + 5x5, C = 21): the fused im2col kernel (tdrn_deform_head) and the project-then-sample pair (tdrn_conv2d_tc +
tdrn_deform_head_sample) for several projection chunk sizes.  CUDA events, 20 iterations after 5 warm-ups.
Not a bench.py number: a development aid for the sampler."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tdrn_b200 import ops

B, C = 32, 21
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
P = 3 * (1600 + 400 + 100 + 25)
loc = torch.empty(B, P, 4, device=dev)
conf = torch.empty(B, P, C, device=dev)
off_p = 0
for hw in [int(v) for v in os.environ.get('LEVELS', '40,20,10,5').split(',')]:
    f = torch.randn(B, hw, hw, 256, generator=g).to(torch.bfloat16).to(dev)
    o1 = (torch.randn(B, hw, hw, 18, generator=g) * 2).to(dev)
    o2 = (torch.randn(B, hw, hw, 50, generator=g) * 2).to(dev)
    w1 = ops.pack_deform_head_weight(torch.randn(12 + 3 * C, 256, 3, 3, generator=g) * 0.02, dev)
    w2 = ops.pack_deform_head_weight(torch.randn(12 + 3 * C, 256, 5, 5, generator=g) * 0.02, dev)
    pc, n_pad = ops.pack_deform_proj_weight(torch.randn(12 + 3 * C, 256, 3, 3, generator=g) * 0.02,
                                            torch.randn(12 + 3 * C, 256, 5, 5, generator=g) * 0.02, dev)
    fl = 2.0 * B * hw * hw * (12 + 3 * C) * 256 * 34
    for mb in (os.environ.get('CHUNKS', '12,24,48,96,100000').split(',')):
        os.environ['TDRN_DEFORM_CHUNK_MB'] = mb
        runp = lambda: ops.deform_head_projected(f, o1, pc, n_pad, C, 3, 1, loc, conf, P, off_p, offsets2=o2, kh2=5, pad2=2, softmax=True)
        for _ in range(5):
            runp()
        ops.prof_begin()
        for _ in range(20):
            runp()
        rec = ops.prof_end()
        tp = sum(r[2] for r in rec if r[0].endswith('project')) / 20
        ts = sum(r[2] for r in rec if r[0].endswith('sample')) / 20
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            runp()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print('projected   %2dx%-2d  chunk %6s MB  %.4f ms  %.1f TFLOP/s nominal   (per-launch events: project %.4f + sample %.4f ms)'
              % (hw, hw, mb, ms, fl / ms / 1e9, tp, ts))
    run = lambda: ops.deform_head(f, o1, w1, C, 1, 3, 1, loc, conf, P, off_p, offsets2=o2, w2_bf16=w2, kh2=5, pad2=2, softmax=True)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('deform head %2dx%-2d  %.4f ms  %.1f TFLOP/s nominal  %.1f G samples/s' % (hw, hw, ms, fl / ms / 1e9, B * hw * hw * 34 * 256 / ms / 1e6))
    off_p += hw * hw * 3
