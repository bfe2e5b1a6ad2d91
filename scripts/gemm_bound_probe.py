"""Probe: what do library kernels reach on the shape of the per-tap projection GEMM (M=51200, N=2720, K=256, bf16 out)
and on a pure 278 MB write / read?  Upper bounds for tdrn_conv2d_tc's resident-weight 1x1 path; development aid."""
import torch
dev = torch.device('cuda')


def t(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


x = torch.randn(51200, 256, device=dev).to(torch.bfloat16)
w = torch.randn(2720, 256, device=dev).to(torch.bfloat16)
y = torch.empty(51200, 2720, device=dev, dtype=torch.bfloat16)
ms = t(lambda: torch.matmul(x, w.t(), out=y))
print('cuBLAS bf16 GEMM 51200x2720x256: %.4f ms  %.1f TFLOP/s  out %.1f MB -> %.2f TB/s written' % (ms, 2 * 51200 * 2720 * 256 / ms / 1e9, y.numel() * 2 / 1e6, y.numel() * 2 / ms / 1e9))
ms = t(lambda: y.zero_())
print('memset 278 MB: %.4f ms  %.2f TB/s' % (ms, y.numel() * 2 / ms / 1e9))
z = torch.empty_like(y)
ms = t(lambda: z.copy_(y))
print('copy 278 MB: %.4f ms  %.2f TB/s (read+write)' % (ms, 2 * y.numel() * 2 / ms / 1e9))
big = torch.empty(1 << 30, device=dev, dtype=torch.uint8)
ms = t(lambda: big.zero_())
print('memset 1 GB: %.4f ms  %.2f TB/s' % (ms, big.numel() / ms / 1e9))
ms = t(lambda: y.sum())
print('read-reduce 278 MB: %.4f ms  %.2f TB/s' % (ms, y.numel() * 2 / ms / 1e9))
