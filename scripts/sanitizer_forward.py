"""One multi-stream b32 forward + Detect of the default workload (bf16) and one b2 forward on the fp32 split path: the
target of compute-sanitizer in scripts/gpu_sanitizer.sh (hand-rolled mbarrier / TMEM pipelines on up to 9 concurrent streams)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
if len(sys.argv) > 2 and sys.argv[2] == 'mobilenet':       # the MobileNet variant (IEEE-half trunk); B >= 64 takes the row-walking depthwise form
    wl = bench.MobileNet()
    wl.setup(torch.device('cuda'), 'bf16')
    x = torch.randn(B, 3, 320, 320, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        out = wl.hot_path(x)
        torch.cuda.synchronize()
    print('mobilenet (half trunk) b%d: %d detections' % (B, int((out[..., 0] > 0).sum())))
    sys.exit(0)
wl = bench.Workload()
wl.setup(torch.device('cuda'), 'bf16')
x = torch.randn(B, 3, 320, 320, generator=torch.Generator().manual_seed(1)).cuda()
with torch.no_grad():
    out = wl.hot_path(x)
    torch.cuda.synchronize()
    print('bf16 b%d: %d detections' % (B, int((out[..., 0] > 0).sum())))
    wl.nets[0].set_precision('fp32')
    out = wl.hot_path(x[:2])
    torch.cuda.synchronize()
    print('fp32 (split tensor-core path) b2: %d detections' % int((out[..., 0] > 0).sum()))
