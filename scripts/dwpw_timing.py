"""Development aid: time tdrn_conv_dwpw against the two-kernel path on the MobileNet-320 b64 shapes (CUDA events, L2 flushed
between launches).  TDRN_DWPW_DEBUG=1/4/5 switch parts of the kernel off (wrong results) to see what bounds it."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from tdrn_b200 import ops

SHAPES = [(64, 160, 160, 32, 64), (64, 80, 80, 128, 128), (64, 80, 80, 256, 256), (64, 40, 40, 512, 512), (64, 20, 20, 1024, 1024)]


def timeit(fn, flush, n=10):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    g = torch.Generator().manual_seed(0)
    bn = lambda c: [torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]
    for (B, H, W, cin, cout) in SHAPES:
        x = torch.randn(B, H, W, cin, generator=g).to(torch.bfloat16).cuda()
        pd = ops.PackedDw(torch.randn(cin, 1, 3, 3, generator=g) * 0.3, bn(cin), 1, 'cuda')
        pc = ops.PackedConv(torch.randn(cout, cin, 1, 1, generator=g) * 0.05, None, bn(cout), 1, 0, 1, device='cuda')
        ops.conv_dwpw(x, pd, pc)
        t_f = timeit(lambda: ops.conv_dwpw(x, pd, pc), flush)
        t_d = timeit(lambda: ops.dwconv3x3(x, pd, relu=True), flush)
        mid = ops.dwconv3x3(x, pd, relu=True)
        t_p = timeit(lambda: ops.conv2d(mid, pc, relu=True, use_tc=True), flush)
        mb = (x.numel() + B * H * W * cout) * 2 / 1e6
        print('%4d -> %4d @%3dx%-3d  fused %.4f ms   dw %.4f + pw %.4f = %.4f ms   (in + out %.0f MB = %.4f ms at 7 TB/s)'
              % (cin, cout, H, W, t_f, t_d, t_p, t_d + t_p, mb, mb / 7e3))


if __name__ == '__main__':
    main()
