"""Pins against the REAL reference native code: the reference's own CUDA sources
(utils/deformconv/deform_conv_cuda_kernel.cu, utils/nms/nms_kernel.cu) are compiled unmodified by
oracle/build_ref.py into oracle/_ref/libtdrn_ref_native.so (built in the container that has /root/reference; the
.so travels to the GPU box) and run here on the B200 next to the oracle restatements and the product kernels.

  * deformable im2col: reference kernel vs oracle (C scalar + torch restatement).  The reference is built
    with nvcc's default -fmad=true (its make.sh:11 passes no flag), so its bilinear blend may contract
    into FMAs; the oracle is contraction-free.  Tolerance: 4 ulp of the largest corner value; the zero /
    non-zero pattern of the border rules (.cu:195-203, :25-37) must match exactly.
  * deformable conv forward: product tdrn_deform_conv_forward (fp32) vs reference im2col + fp32 GEMM, 1e-4.
  * NMS: reference GPU `_nms` (suppress when IoU > thresh) vs product tdrn_nms (IoU >= thresh, the CPU rule
    Detect uses): identical keep lists whenever no pair sits exactly on the threshold; and product tdrn_nms vs the
    reference's own compiled Cython cpu_nms (oracle/build_ref_nms.py), pair exactly on the threshold included.
"""
import numpy as np
import pytest
import torch

from oracle import ref_native, c_oracle, deform_conv_ref, nms_ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_native.available(), reason='oracle/_ref not built (needs /root/reference)')]

CASES = [  # C, H, W, k, stride, pad, dil, dg, offset sigma
    (8, 9, 11, 3, 1, 1, 1, 1, 2.0),
    (16, 10, 10, 3, 1, 1, 1, 2, 3.0),
    (64, 20, 20, 5, 1, 2, 1, 1, 2.0),
    (32, 13, 17, 3, 2, 1, 1, 4, 1.5),
    (16, 12, 12, 3, 1, 2, 2, 8, 2.5),
    (6, 32, 32, 3, 1, 1, 1, 2, 0.0),      # the shape family of utils/deformconv/test.py (6 ch, dg 2), zero offsets
]


def _case(c, h, w, k, s, p, d, dg, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    ho = (h + 2 * p - (d * (k - 1) + 1)) // s + 1
    wo = (w + 2 * p - (d * (k - 1) + 1)) // s + 1
    inp = torch.randn(c, h, w, generator=g)
    off = torch.randn(dg * 2 * k * k, ho, wo, generator=g) * sigma
    # exact-integer and exact-border offsets exercise the `h_low >= H-1` and validity branches
    off.view(-1)[::7] = torch.round(off.view(-1)[::7])
    return inp, off


@pytest.mark.parametrize('case', CASES)
def test_reference_im2col_kernel_matches_oracle(case):
    c, h, w, k, s, p, d, dg, sigma = case
    inp, off = _case(*case, seed=3)
    ref = ref_native.deformable_im2col(inp.cuda(), off.cuda(), k, k, s, p, d, dg).cpu()
    ora = deform_conv_ref.deform_im2col(inp, off, k, k, s, p, d, dg)
    assert ref.shape == ora.shape
    tol = 4 * np.finfo(np.float32).eps * float(inp.abs().max())
    assert float((ref - ora).abs().max()) <= tol
    # border semantics: samples the reference zeroes are exactly the samples the oracle zeroes
    assert torch.equal(ref == 0, ora == 0)
    # the C restatement agrees with the torch restatement bit for bit, hence with the reference to `tol`
    wone = torch.zeros(1, c, k, k)
    wone[0, 0, 0, 0] = 1.0                       # picks column row 0 out of the C oracle's conv
    got = c_oracle.deform_conv_forward(inp[None].numpy(), off[None].numpy(), wone.numpy(), s, p, d, dg)
    assert np.array_equal(got.reshape(-1), ora[0].numpy())


@pytest.mark.parametrize('case', CASES)
def test_product_deform_conv_matches_reference_kernel(case):
    from tdrn_b200.model.networks import conv_offset2d
    c, h, w, k, s, p, d, dg, sigma = case
    inp, off = _case(*case, seed=5)
    B, cout = 3, 21
    g = torch.Generator().manual_seed(9)
    x = torch.stack([inp * (i + 1) for i in range(B)]).cuda()
    o = torch.stack([off.roll(i, 1) for i in range(B)]).cuda()
    wt = (torch.randn(cout, c, k, k, generator=g) / (c * k * k) ** 0.5).cuda()
    ref = ref_native.deform_conv_forward(x, o, wt, s, p, d, dg)
    got = conv_offset2d(x, o, wt, s, p, d, dg)
    assert got.shape == ref.shape
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err


def _boxes(n, seed, spread=300.0):
    rng = np.random.RandomState(seed)
    xy = rng.uniform(0, spread, (n, 2))
    wh = rng.uniform(8, 90, (n, 2))
    s = rng.permutation(n).astype(np.float32) / n + 0.001          # distinct scores: no tie-order ambiguity
    return np.hstack([xy, xy + wh, s[:, None]]).astype(np.float32)


@pytest.mark.parametrize('n,thresh', [(1, 0.45), (63, 0.45), (64, 0.3), (65, 0.5), (1000, 0.45), (6375, 0.45)])
def test_product_nms_matches_reference_gpu_kernel(n, thresh):
    from tdrn_b200.utils.nms_wrapper import nms
    dets = _boxes(n, n)
    ref = ref_native.gpu_nms(dets, thresh)
    got = nms(dets, thresh)
    ora = nms_ref.cpu_nms(dets, thresh)
    assert [int(i) for i in ref] == [int(i) for i in got] == [int(i) for i in ora]
    # a pair exactly on the threshold: the reference's GPU kernel keeps it (`>`, nms_kernel.cu:71), and so does nms() with the
    # default force_cpu=False; the CPU rule (`>=`) drops it
    edge = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 130, 0.7]], np.float32)      # IoU(0, 1) == 0.5
    assert [int(i) for i in ref_native.gpu_nms(edge, 0.5)] == nms(edge, 0.5) == [0, 1, 2]
    assert nms(edge, 0.5, force_cpu=True) == [0, 2]


@pytest.mark.parametrize('n,thresh', [(1, 0.45), (2, 0.45), (257, 0.3), (1000, 0.45), (3000, 0.6), (6375, 0.45)])
def test_product_nms_matches_reference_cython_nms(n, thresh):
    """tdrn_nms (through utils.nms_wrapper.nms) == the reference's own compiled cpu_nms.pyx -- the function Detect calls --
    built by oracle/build_ref_nms.py; includes a pair exactly on the threshold (`>=` suppresses, cpu_nms.pyx:65)."""
    from oracle import build_ref_nms
    mod = build_ref_nms.load()
    if mod is None:
        pytest.skip('oracle/_ref/cpu_nms*.so not built (needs /root/reference + Cython)')
    from tdrn_b200.utils.nms_wrapper import nms
    dets = _boxes(n, 100 + n)
    assert [int(i) for i in nms(dets, thresh, force_cpu=True)] == [int(i) for i in mod.cpu_nms(dets, thresh)]
    edge = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 19, 0.8], [100, 100, 120, 130, 0.7]], np.float32)   # IoU(0, 1) == 0.5
    assert [int(i) for i in nms(edge, 0.5, force_cpu=True)] == [int(i) for i in mod.cpu_nms(edge, 0.5)] == [0, 2]


def test_deform_head_vs_the_reference_kernels_speed_and_values():
    """DeformConv head-to-head on the same B200 (BASELINE.json: "DeformConv % roofline"): the ODM head of pyramid level 0
    at the bench shape (b32, 256 ch, 40x40, 3x3 loc/conf + 5x5 multihead, VOC-21) through
      (a) the reference's own path: its unmodified CUDA im2col kernel per sample + fp32 cuBLAS GEMM, for loc, conf, loc_2,
          conf_2 (deform_conv_cuda.c:157-193; 256 launches, replayed as ONE CUDA graph so that launch overhead is not held
          against it), softmax by torch;
      (b) tdrn_b200: per-tap projection GEMM + sampler (2 launches).
    Values agree within the bf16 bar; the timings are written to gpurun_out/deform_vs_reference.txt."""
    import os
    from tdrn_b200 import ops
    B, C, H, W, ncls = 32, 256, 40, 40, 21
    g = torch.Generator().manual_seed(11)
    bf = lambda t: t.to(torch.bfloat16).float()
    x = bf(torch.randn(B, C, H, W, generator=g)).cuda()
    off1 = (torch.randn(B, 18, H, W, generator=g) * 1.5).cuda()
    off2 = (torch.randn(B, 50, H, W, generator=g) * 1.5).cuda()
    wl, wc = bf(torch.randn(12, C, 3, 3, generator=g) * 0.03), bf(torch.randn(3 * ncls, C, 3, 3, generator=g) * 0.03)
    wl2, wc2 = bf(torch.randn(12, C, 5, 5, generator=g) * 0.02), bf(torch.randn(3 * ncls, C, 5, 5, generator=g) * 0.02)
    wl_c, wc_c, wl2_c, wc2_c = wl.cuda(), wc.cuda(), wl2.cuda(), wc2.cuda()

    def reference_head():
        loc = ref_native.deform_conv_forward(x, off1, wl_c, 1, 1, 1, 1) + ref_native.deform_conv_forward(x, off2, wl2_c, 1, 2, 1, 1)
        conf = ref_native.deform_conv_forward(x, off1, wc_c, 1, 1, 1, 1) + ref_native.deform_conv_forward(x, off2, wc2_c, 1, 2, 1, 1)
        loc = loc.permute(0, 2, 3, 1).reshape(B, -1, 4)                                # dualrefinedet_vggbn.py:186-189
        conf = torch.softmax(conf.permute(0, 2, 3, 1).reshape(-1, ncls), 1)            # :196
        return loc, conf

    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    xb = nhwc(x).to(torch.bfloat16)
    o1, o2 = nhwc(off1), nhwc(off2)
    pc, n_pad = ops.pack_deform_proj_weight(torch.cat([wl, wc], 0), torch.cat([wl2, wc2], 0))
    P = H * W * 3
    loc_o = torch.zeros(B, P, 4, device='cuda')
    conf_o = torch.zeros(B, P, ncls, device='cuda')

    def ours():
        ops.deform_head_projected(xb, o1, pc, n_pad, ncls, 3, 1, loc_o, conf_o, P, 0, offsets2=o2, kh2=5, pad2=2, softmax=True)

    def timed(fn, graph):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st), torch.no_grad():
            for _ in range(2):
                fn()
            st.synchronize()
            run = fn
            if graph:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    fn()
                run = gr.replay
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10):
                run()
            e1.record(st)
            st.synchronize()
        return e0.elapsed_time(e1) / 10

    loc_r, conf_r = reference_head()
    ours()
    torch.cuda.synchronize()
    assert float((loc_o - loc_r).abs().max() / loc_r.abs().max()) < 2e-2
    assert float((conf_o.view(-1, ncls) - conf_r).abs().max() / conf_r.abs().max()) < 2e-2
    ms_ref = timed(reference_head, graph=True)
    ms_ours = timed(ours, graph=True)
    flops = 2.0 * B * H * W * 75 * C * 34
    line = ('DeformConv head, level 0 of DualRefineDet-VGGBN-320 b32 (multihead, VOC-21), one B200:\n'
            '  reference CUDA im2col kernel + fp32 cuBLAS GEMM (graph replay): %.3f ms  (%.1f TFLOP/s nominal)\n'
            '  tdrn_b200 projection GEMM + sampler (graph replay):             %.3f ms  (%.1f TFLOP/s nominal)\n'
            '  speed-up %.1fx\n' % (ms_ref, flops / ms_ref / 1e9, ms_ours, flops / ms_ours / 1e9, ms_ref / ms_ours))
    os.makedirs('gpurun_out', exist_ok=True)
    open('gpurun_out/deform_vs_reference.txt', 'w').write(line)
    print(line)
    assert ms_ours < ms_ref
