"""Host-side executor shared by the detector modules.

The reference forward passes (model/dualrefinedet_vggbn.py:119-206,
model/dualrefinedet_mobilenet.py:127-199, model/refinedet_vgg.py:109-219,
model/ssd4scale_vgg.py:71-135) are re-expressed as sequences of C-ABI operator calls on NHWC
activations.  BatchNorm is folded into the preceding conv at pack time, ReLU / bias / residual add are
conv epilogues, permute(0,2,3,1)+view+cat of the heads disappears because NHWC heads write straight
into the flattened [B,P,4] / [B,P,C] outputs.

precision: 'fp32' -> fp32 activations, SIMT fp32 kernels (1e-4 path)
           'bf16' -> bf16 activations, tcgen05 implicit-GEMM convs + fused deformable head (2e-2 path)
"""
import os

import torch

from .. import ops

VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512]
NUM_BOX = 3


def level_sizes(size):
    """Feature-map sides of the four pyramid levels (stride 8/16/32/64; pool3 is ceil_mode)."""
    s = size
    s = s // 2            # pool1
    s = s // 2            # pool2
    s = (s + 1) // 2      # pool3 'C'
    l0 = s
    l1 = l0 // 2          # pool4
    l2 = l1 // 2          # pool5
    l3 = (l2 + 2 - 3) // 2 + 1   # extras 3x3 s2 p1
    return [l0, l1, l2, l3]


def _mark_split(t, kw):
    if kw.get('split_out'):
        t._tdrn_is_split = True
    return t


class Engine(object):
    def __init__(self, module, precision):
        if precision not in ('fp32', 'bf16'):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self.act = torch.bfloat16 if precision == 'bf16' else torch.float32
        self.use_tc = precision == 'bf16' and os.environ.get('TDRN_DISABLE_TC', '0') != '1'
        # fp32 path: dense convs with Cin % 64 == 0 run on the tensor cores too, as three bf16 products per fp32 product
        # (x = hi + lo, w = hi + lo; hi*hi + hi*lo + lo*hi with fp32 accumulation: 16 mantissa bits per operand, 1-2e-5
        # relative on the detector outputs against the 1e-4 bar); TDRN_FP32_SIMT=1 keeps every conv on the CUDA cores
        # (the MobileNet trunks amplify per-layer error ~3x more than VGG -- DESIGN.md section 5 -- and land at 1.4-2.9e-4 with the
        # split products, above the 1e-4 bar: those modules set fp32_tensor_cores = False and keep the CUDA-core fp32 convs)
        self.use_x3 = (precision == 'fp32' and os.environ.get('TDRN_FP32_SIMT', '0') != '1'
                       and getattr(module, 'fp32_tensor_cores', True))
        self.sd = {k: v for k, v in module.state_dict().items()}
        p = next(module.parameters())
        if not p.is_cuda:
            raise NotImplementedError('tdrn_b200 runs on CUDA only: call net.to("cuda") first (no CPU fallback)')
        self.device = p.device
        self.pk = {}
        self.multi_stream = os.environ.get('TDRN_SINGLE_STREAM', '0') != '1'
        self._side = []
        self._rr = 0
        self.early_fork = os.environ.get('TDRN_EARLY_FORK', '0') != '0'   # measured: forking under conv5_x steals SMs from the critical path
        # MobileNet trunks in the 16-bit mode: IEEE half instead of bf16 for the trunk's activations and weights (same tensor-core
        # rate, 11 significand bits instead of 8).  With bf16 storage the 27 stacked layers land at ~3e-2 on the detector outputs,
        # above the 2e-2 bar; the sources leave the trunk as bf16 for the ARM heads / TCB / deformable heads.
        self.half_trunk = self.use_tc and os.environ.get('TDRN_MOBILE_BF16', '0') != '1'

    # ---- weight packing -------------------------------------------------------------------------
    def _bn(self, name):
        return (self.sd[name + '.weight'], self.sd[name + '.bias'], self.sd[name + '.running_mean'],
                self.sd[name + '.running_var'])

    def packed(self, name, stride=1, pad=0, dil=1, bn=None, deconv=False):
        pc = self.pk.get(name)
        if pc is None:
            pc = ops.PackedConv(self.sd[name + '.weight'], self.sd.get(name + '.bias'),
                                self._bn(bn) if bn else None, stride, pad, dil, deconv, self.device,
                                want_bf16=self.precision == 'bf16', want_x3=self.use_x3)
            self.pk[name] = pc
        return pc

    def packed_dw(self, name, bn, stride):
        pd = self.pk.get(name)
        if pd is None:
            pd = ops.PackedDw(self.sd[name + '.weight'], self._bn(bn), stride, self.device)
            self.pk[name] = pd
        return pd

    def vec(self, name):
        v = self.pk.get(name)
        if v is None:
            v = self.sd[name].detach().float().contiguous().to(self.device)
            self.pk[name] = v
        return v

    # ---- stream-level concurrency -----------------------------------------------------------------
    def parallel(self, fns):
        """Run independent closures on side streams (fork after the current stream, join back into it).
        The four pyramid levels' heads and the TCB branches are independent and each of them is far too
        small to fill 148 SMs (7..100 CTAs), so they are issued as parallel branches; under CUDA-graph
        capture the event dependencies become parallel graph branches.  Returns the closures' results."""
        if len(fns) <= 1 or not self.multi_stream:
            return [f() for f in fns]
        main = torch.cuda.current_stream()
        while len(self._side) < len(fns) - 1:
            self._side.append(torch.cuda.Stream(self.device))
        fork = torch.cuda.Event()
        fork.record(main)
        results = [None] * len(fns)
        joins = []
        for i, f in enumerate(fns[1:]):
            st = self._side[i]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                results[i + 1] = f()
                ev = torch.cuda.Event()
                ev.record(st)
            joins.append(ev)
            for t in _tensors(results[i + 1]):
                t.record_stream(main)            # allocated on the side stream, consumed on the main one
        results[0] = fns[0]()
        for ev in joins:
            main.wait_event(ev)
        return results

    def spawn(self, fn):
        """Start ``fn`` on a side stream forked from the current point of the current stream (so it may begin
        while later trunk layers are still being issued on the main stream).  -> handle for ``join``."""
        if not self.multi_stream:
            return (None, fn())
        main = torch.cuda.current_stream()
        if len(self._side) < 8:
            self._side.append(torch.cuda.Stream(self.device))
        st = self._side[self._rr % len(self._side)]
        self._rr += 1
        fork = torch.cuda.Event()
        fork.record(main)
        st.wait_event(fork)
        with torch.cuda.stream(st):
            res = fn()
            done = torch.cuda.Event()
            done.record(st)
        for t in _tensors(res):
            t.record_stream(main)
        return (done, res)

    def join(self, handle):
        done, res = handle
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
        return res

    # ---- operators ------------------------------------------------------------------------------
    def conv(self, name, x, stride=1, pad=0, dil=1, bn=None, relu=False, deconv=False, **kw):
        pc = self.packed(name, stride, pad, dil, bn, deconv)
        use_tc = (self.use_tc and pc.w_bf16 is not None and x.dtype in (torch.bfloat16, torch.float16)
                  and not kw.get('dg') and kw.get('in_shape') is None and (stride in (1, 2) or deconv))
        ceil_mode = kw.pop('ceil_mode', False)
        to_split = kw.pop('to_split', False)
        is_split = getattr(x, '_tdrn_is_split', False)
        if is_split and not (self.use_x3 and pc.w_x3 is not None):
            raise RuntimeError('%s: a split (hi | lo) activation can only feed a split-precision tensor-core conv' % name)
        if (self.use_x3 and pc.w_x3 is not None and (is_split or x.dtype == torch.float32) and not kw.get('dg')
                and kw.get('in_shape') is None and (stride in (1, 2) or deconv)):
            # one split per activation tensor, shared by all its consumers -- which may sit on different branch streams (an ARM
            # source feeds its ARM head and its TCB branch): the split is made on the first consumer's stream and carries an
            # event the others wait for.  Layers whose output only feeds other split-precision convs (to_split=True) write the
            # (hi | lo) operand themselves: no fp32 copy of it, no split kernel.
            if to_split and not deconv and kw.get('residual') is None and kw.get('out') is None and pc.cout % 16 == 0:
                kw['split_out'] = pc.cout
            xs = self.split_of(x)
            kw.setdefault('out_dtype', torch.float32)
            if kw.pop('pool', False):
                H, W = x.shape[1], x.shape[2]
                if stride == 1 and 2 * pad == dil * (pc.kh - 1) and W % 16 == 0 and H % 8 == 0:
                    return _mark_split(ops.conv2d(xs, pc, relu=relu, use_tc=True, pool=True, split3=True, **kw), kw)
                kw.pop('split_out', None)
                return ops.maxpool2x2(ops.conv2d(xs, pc, relu=relu, use_tc=True, split3=True, **kw), ceil_mode)
            return _mark_split(ops.conv2d(xs, pc, relu=relu, use_tc=True, split3=True, **kw), kw)
        if kw.pop('pool', False):
            # MaxPool2d(2,2) after the conv: fused into the tcgen05 epilogue when the map tiles as 16x8 boxes
            H, W = x.shape[1], x.shape[2]
            if use_tc and stride == 1 and 2 * pad == dil * (pc.kh - 1) and W % 16 == 0 and H % 8 == 0:
                return ops.conv2d(x, pc, relu=relu, use_tc=True, pool=True, **kw)
            return ops.maxpool2x2(ops.conv2d(x, pc, relu=relu, use_tc=use_tc, **kw), ceil_mode)
        return ops.conv2d(x, pc, relu=relu, use_tc=use_tc, **kw)

    def split_of(self, x):
        """The (hi | lo) bf16 operand of fp32 activation ``x`` (made once, cached on the tensor with the event other streams wait for)."""
        if getattr(x, '_tdrn_is_split', False):
            return x
        cur = torch.cuda.current_stream()
        ent = getattr(x, '_tdrn_split', None)
        if ent is None:
            xs = ops.split_bf16(x)
            ev = torch.cuda.Event()
            ev.record(cur)
            try:
                x._tdrn_split = (xs, ev, cur)
            except Exception:
                pass
            return xs
        xs, ev, st = ent
        if st != cur:
            cur.wait_event(ev)
            xs.record_stream(cur)
        return xs

    def conv_first(self, name, x_nchw, stride, bn, to_split=False, out_dtype=None):
        pc = self.packed(name, stride, 1, 1, bn)
        if out_dtype is not None:
            return ops.conv_first(x_nchw, pc, True, out_dtype)
        if to_split and self.use_x3 and ops.conv_first_split_ok(x_nchw, pc):
            # fp32 path, VGG conv1_1: split-precision tensor-core stem writing conv1_2's (hi | lo) operand
            y = ops.conv_first(x_nchw, pc, True, self.act, split=True)
            y._tdrn_is_split = True
            return y
        return ops.conv_first(x_nchw, pc, True, self.act)

    # ---- VGG trunk: vgg() model/networks.py:136-163 + extras, forward :130-153 --------------------
    def vgg_trunk(self, x_nchw, bn, with_extras=True, on_source=None):
        """``on_source(k, tensor)`` is called the moment ARM source k exists (lets the caller fork its heads early)."""
        split43, split53 = (23, 33)[bn], (30, 43)[bn]
        sources = []
        idx, x = 0, None
        step = 3 if bn else 2

        def conv_block(i, x, pad=1, dil=1, pool=None, to_split=False):
            # to_split: the output feeds only the next trunk conv (fp32 path: written as the (hi | lo) operand, see conv())
            bnn = 'backbone.%d' % (i + 1) if bn else None
            if x is None:
                return self.conv_first('backbone.%d' % i, x_nchw, 1, bnn, to_split=to_split)
            if pool is not None:
                return self.conv('backbone.%d' % i, x, 1, pad, dil, bn=bnn, relu=True, pool=True, ceil_mode=pool, to_split=to_split)
            return self.conv('backbone.%d' % i, x, 1, pad, dil, bn=bnn, relu=True, to_split=to_split)

        n_cfg = len(VGG_CFG)
        ci = 0
        if (self.use_tc and self.act == torch.bfloat16 and VGG_CFG[:3] == [64, 64, 'M']
                and os.environ.get('TDRN_NO_STEM_PAIR') is None):
            # conv1_1 -> conv1_2 -> pool1 in one kernel: conv1_1's 419 MB (b32) output never reaches HBM
            pc1 = self.packed('backbone.0', 1, 1, 1, 'backbone.1' if bn else None)
            pc2 = self.packed('backbone.%d' % step, 1, 1, 1, 'backbone.%d' % (step + 1) if bn else None)
            y = ops.conv_stem_pair(x_nchw, pc1, pc2, relu=True, pool=True)
            if y is not None:
                x, idx, ci = y, 2 * step + 1, 3
        while ci < n_cfg:
            v = VGG_CFG[ci]
            if idx == split43:
                if v == 'M':                              # conv4_3 feeds L2Norm_4_3 and pool4: one pass over x
                    s0, x = ops.l2norm_pool(x, self.vec('L2Norm_4_3.weight'))
                    sources.append(s0)
                    if on_source:
                        on_source(0, s0)
                    idx += 1
                    ci += 1
                    continue
                sources.append(ops.l2norm(x, self.vec('L2Norm_4_3.weight')))
                if on_source:
                    on_source(0, sources[0])
            if v == 'M' or v == 'C':
                x = ops.maxpool2x2(x, ceil_mode=(v == 'C'))
                idx += 1
            else:
                nxt = VGG_CFG[ci + 1] if ci + 1 < n_cfg else None
                # conv followed by a pool whose input nobody else needs (not conv4_3 -> L2Norm): fuse the pool
                if x is not None and nxt in ('M', 'C') and idx + step != split43 and (nxt == 'M' or x.shape[1] % 2 == 0):
                    x = conv_block(idx, x, pool=(nxt == 'C'), to_split=True)
                    idx += step + 1
                    ci += 1
                else:
                    # conv4_3 / conv5_3 feed an L2Norm, a conv before an unfused pool feeds the pool kernel: those stay fp32
                    x = conv_block(idx, x, to_split=isinstance(nxt, int) and idx + step not in (split43, split53))
                    idx += step
            ci += 1
        assert idx == split53
        s1, x = ops.l2norm_pool(x, self.vec('L2Norm_5_3.weight'))   # conv5_3 -> L2Norm_5_3 and pool5 (pool5_ds=True)
        sources.append(s1)
        if on_source:
            on_source(1, s1)
        idx += 1
        x = conv_block(idx, x, 6, 6, to_split=True)       # conv6: 3x3 dilation 6
        idx += step
        x = conv_block(idx, x, 0, 1)                      # conv7: 1x1
        sources.append(x)
        if on_source:
            on_source(2, x)
        if with_extras:
            if bn:
                x = self.conv('extras.0', x, bn='extras.1', relu=True, to_split=True)
                x = self.conv('extras.3', x, 2, 1, bn='extras.4', relu=True)
            else:
                x = self.conv('extras.0', x, relu=True, to_split=True)
                x = self.conv('extras.2', x, 2, 1, relu=True)
            sources.append(x)
            if on_source:
                on_source(3, x)
        return sources

    # ---- TCB / FPN: dualrefinedet_vggbn.py:166-179 -------------------------------------------------
    def fpn(self, arm_sources):
        x = self.last_trans(arm_sources[3])
        odm = [x]
        return self.fpn_topdown(x, odm, [self.trans_branch(arm_sources[k], k) for k in range(3)])

    def last_trans(self, src3):
        x = self.conv('last_layer_trans.0', src3, 1, 1, relu=True, to_split=True)
        x = self.conv('last_layer_trans.2', x, 1, 1, to_split=True)
        return self.conv('last_layer_trans.3', x, 1, 1)

    def trans_branch(self, src, k):
        t = self.conv('trans_layers.%d.0' % k, src, 1, 1, relu=True, to_split=True)
        return self.conv('trans_layers.%d.2' % k, t, 1, 1)

    def fpn_topdown(self, x, odm, trans):
        for k in range(3):
            t = trans[2 - k]
            # relu(up(x) + t): ConvTranspose2d k2 s2 as a pixel-shuffled GEMM with residual epilogue
            u = self.conv('up_layers.%d' % k, x, deconv=True, relu=True, residual=t)
            x = self.conv('latent_layers.%d' % k, u, 1, 1, relu=True)
            odm.append(x)
        odm.reverse()
        return odm

    # ---- heads ----------------------------------------------------------------------------------
    def head_into(self, name, x, flat, per_prior, prior_off, P, pad=1, residual=False, offsets=None, dg=0):
        """conv head writing NHWC-flattened rows into flat [B, P*per_prior] at prior offset."""
        B = x.shape[0]
        cout = NUM_BOX * per_prior
        view = flat.view(B, P * per_prior)[:, prior_off * per_prior:]
        return self.conv(name, x, 1, pad, out=view, out_sb=P * per_prior, out_sp=cout,
                         residual=view if residual else None, offsets=offsets, dg=dg)

    def offset_maps(self, arm_loc, sizes, lv_off, multihead, want_nchw):
        """`offset.k` / `offset2.k` 1x1 convs of every level on the flattened ARM regression, one launch
        (dualrefinedet_vggbn.py:160-164).  -> offsets NHWC, offsets2 NHWC (or []), offsets NCHW (or None)."""
        key = 'offset_w.%d' % int(multihead)
        w = self.pk.get(key)
        if w is None:
            n = len(sizes)
            f = lambda name: self.sd[name].detach().float().reshape(self.sd[name].shape[0], -1).contiguous().to(self.device)
            b = lambda name: (self.sd[name].detach().float().contiguous().to(self.device) if name in self.sd else None)
            w = ([f('offset.%d.weight' % k) for k in range(n)], [b('offset.%d.bias' % k) for k in range(n)],
                 [f('offset2.%d.weight' % k) for k in range(n)] if multihead else None,
                 [b('offset2.%d.bias' % k) for k in range(n)] if multihead else None)
            self.pk[key] = w
        b1 = w[1] if all(t is not None for t in w[1]) else None
        b2 = w[3] if (multihead and all(t is not None for t in w[3])) else None
        return ops.offset_convs(arm_loc, sizes, lv_off, w[0], b1, w[2], b2, want_nchw=want_nchw)

    def arm_heads(self, arm_sources, P, lv_off, multihead, with_offsets=True):
        """arm_loc + 1x1 offset convs: dualrefinedet_vggbn.py:154-165."""
        B = arm_sources[0].shape[0]
        arm_loc = torch.empty(B, P, 4, dtype=torch.float32, device=self.device)

        def level(k):
            self.head_into('arm_loc.%d' % k, arm_sources[k], arm_loc, 4, lv_off[k], P)

        self.parallel([(lambda k=k: level(k)) for k in range(len(arm_sources))])
        if not with_offsets:
            return arm_loc, [], []
        sizes = [(a.shape[1], a.shape[2]) for a in arm_sources]
        offs, offs2, _ = self.offset_maps(arm_loc, sizes, lv_off, multihead, want_nchw=False)
        return arm_loc, offs, offs2

    def arm_and_tcb(self, arm_sources, P, lv_off, multihead):
        """ARM heads (+offset convs) of the four levels, the three TCB transfer branches and the top
        last_layer_trans chain are mutually independent: one fork/join."""
        B = arm_sources[0].shape[0]
        arm_loc = torch.empty(B, P, 4, dtype=torch.float32, device=self.device)

        def level(k):
            self.head_into('arm_loc.%d' % k, arm_sources[k], arm_loc, 4, lv_off[k], P)

        fns = [lambda: self.last_trans(arm_sources[3])]
        fns += [(lambda k=k: self.trans_branch(arm_sources[k], k)) for k in range(3)]
        fns += [(lambda k=k: level(k)) for k in range(4)]
        res = self.parallel(fns)
        x, trans = res[0], res[1:4]
        offs, offs2, _ = self.offset_maps(arm_loc, [(a.shape[1], a.shape[2]) for a in arm_sources], lv_off, multihead, False)
        odm = self.fpn_topdown(x, [x], trans)
        return arm_loc, offs, offs2, odm

    def trunk_arm_tcb(self, x_nchw, bn, size, multihead, want_nchw_offsets=True):
        """VGG trunk with the ARM heads (+1x1 offset convs) and the TCB transfer branches forked the moment their
        source exists (dualrefinedet_vggbn.py:130-179): the 40x40 / 20x20 branches run under conv5_x .. extras
        instead of after them, and the FPN top-down chain only waits for the branch it consumes.
        -> arm_loc, offsets (NHWC), offsets2 (NHWC), offsets NCHW (or None), odm sources."""
        B = x_nchw.shape[0]
        sizes = level_sizes(size)
        lv_off, P = [], 0
        for sd in sizes:
            lv_off.append(P)
            P += sd * sd * NUM_BOX
        arm_loc = torch.empty(B, P, 4, dtype=torch.float32, device=self.device)
        handles = {}

        def level(k, a):
            self.head_into('arm_loc.%d' % k, a, arm_loc, 4, lv_off[k], P)

        def on_source(k, a):
            assert a.shape[1] == sizes[k], (a.shape, sizes)
            if k < 3:
                handles['trans', k] = self.spawn(lambda: self.trans_branch(a, k))
            else:
                handles['last'] = self.spawn(lambda: self.last_trans(a))
            handles['arm', k] = self.spawn(lambda: level(k, a))

        if self.early_fork:
            self.vgg_trunk(x_nchw, bn, on_source=on_source)
        else:
            for k, a in enumerate(self.vgg_trunk(x_nchw, bn)):
                on_source(k, a)
        def offsets_branch():
            # the 1x1 offset convs of all four levels (+ the NCHW maps the reference returns) in one launch, once every ARM head is
            # in -- on a side stream, under the FPN top-down chain
            for k in range(4):
                self.join(handles['arm', k])
            return self.offset_maps(arm_loc, [(sd, sd) for sd in sizes], lv_off, multihead, want_nchw_offsets)

        h_off = self.spawn(offsets_branch)
        x = self.join(handles['last'])
        odm = [x]
        for k in range(3):
            t = self.join(handles['trans', 2 - k])
            u = self.conv('up_layers.%d' % k, x, deconv=True, relu=True, residual=t)
            x = self.conv('latent_layers.%d' % k, u, 1, 1, relu=True)
            odm.append(x)
        odm.reverse()
        offs, offs2, offs_nchw = self.join(h_off)
        return arm_loc, offs, offs2, offs_nchw, odm, P, lv_off

    def deform_heads(self, feats, offs, offs2, P, lv_off, num_classes, dg, multihead, loc_name='odm_loc',
                     conf_name='odm_conf', softmax=True):
        """Deformable loc/conf heads (+5x5 multihead) and softmax: dualrefinedet_vggbn.py:180-197."""
        B = feats[0].shape[0]
        loc = torch.empty(B, P, 4, dtype=torch.float32, device=self.device)
        conf = torch.empty(B, P, num_classes, dtype=torch.float32, device=self.device)
        # tdrn_deform_head (csrc/deform_tc.cu) needs whole 64-channel blocks per deformable group and the fused loc||conf
        # accumulator (12 + 3C columns, rounded up to 16) inside one 256-column TMEM tile; other shapes (e.g. def_groups = 8 on
        # the 256-channel ODM maps, or more than 81 classes) take the SIMT deformable heads
        n_out16 = (12 + 3 * num_classes + 15) // 16 * 16
        fused = (self.use_tc and feats[0].dtype == torch.bfloat16 and n_out16 <= 256
                 and all(f.shape[3] % 64 == 0 and (f.shape[3] // max(dg, 1)) % 64 == 0 for f in feats))

        # narrow head, one deformable group: project per tap on the tensor cores first, then sample the projections
        # (3.4x fewer bilinear samples at VOC-21); wide heads / dg > 1 sample into the fused im2col tile instead
        n_out8 = (12 + 3 * num_classes + 7) // 8 * 8
        projected = (fused and dg == 1 and os.environ.get('TDRN_DEFORM_PATH', 'project') != 'im2col'
                     and all(2 * n_out8 <= f.shape[3] for f in feats))

        # fp32 path: the same "project, then sample" head with the projection GEMM in split precision and the projections kept
        # as (hi | lo) bf16 pairs (both halves sampled, added in fp32): replaces the CUDA-core deformable convs (24 ms of a
        # 28 ms fp32 step at b32) when the head is narrow enough (12 + 3C <= 128)
        projected_x3 = (self.use_x3 and dg == 1 and feats[0].dtype == torch.float32 and n_out16 <= 128
                        and all(f.shape[3] % 64 == 0 for f in feats) and os.environ.get('TDRN_DEFORM_PATH', 'project') == 'project')

        def level(k):
            f = feats[k]
            if projected or projected_x3:
                pc, n_pad = self.projected_head_weight(loc_name, conf_name, k, multihead, x3=projected_x3)
                ops.deform_head_projected(f, offs[k], pc, n_pad, num_classes, 3, 1, loc, conf, P, lv_off[k],
                                          offsets2=offs2[k] if multihead else None,
                                          kh2=5 if multihead else 0, pad2=2 if multihead else 0, softmax=softmax,
                                          split=projected_x3)
            elif fused:
                w1 = self.fused_head_weight(loc_name, conf_name, k)
                w2 = self.fused_head_weight(loc_name + '_2', conf_name + '_2', k) if multihead else None
                ops.deform_head(f, offs[k], w1, num_classes, dg, 3, 1, loc, conf, P, lv_off[k],
                                offsets2=offs2[k] if multihead else None, w2_bf16=w2,
                                kh2=5 if multihead else 0, pad2=2 if multihead else 0, softmax=softmax)
            else:  # noqa: E101
                self.head_into('%s.%d' % (loc_name, k), f, loc, 4, lv_off[k], P, 1, offsets=offs[k], dg=dg)
                self.head_into('%s.%d' % (conf_name, k), f, conf, num_classes, lv_off[k], P, 1, offsets=offs[k], dg=dg)
                if multihead:
                    self.head_into('%s_2.%d' % (loc_name, k), f, loc, 4, lv_off[k], P, 2, True, offs2[k], dg)
                    self.head_into('%s_2.%d' % (conf_name, k), f, conf, num_classes, lv_off[k], P, 2, True, offs2[k], dg)

        grouped = ((projected or projected_x3) and len(feats) <= 6 and os.environ.get('TDRN_DEFORM_GROUP', '1') != '0'
                   and all(f.shape[0] * f.shape[1] * f.shape[2] * 34 * 2 * (2 * n_out16 if projected_x3 else n_out8) <= ops._deform_chunk_bytes()
                           for f in feats))
        if grouped:
            # projections of the levels on parallel branches, then ONE sampler launch over all levels (the small levels are
            # launch / tail bound on their own)
            def project(k):
                pc, n_pad = self.projected_head_weight(loc_name, conf_name, k, multihead, x3=projected_x3)
                return ops.deform_project(feats[k], pc, n_pad, 3, 5 if multihead else 0, num_classes, split=projected_x3,
                                          xs=self.split_of(feats[k]) if projected_x3 else None)

            ys = self.parallel([(lambda k=k: project(k)) for k in range(len(feats))])
            n_pad = self.projected_head_weight(loc_name, conf_name, 0, multihead, x3=projected_x3)[1]
            ops.deform_sample_group(ys, [tuple(f.shape) for f in feats], n_pad, num_classes, 3, 1, offs, loc, conf, P, lv_off,
                                    offsets2=offs2 if multihead else None, kh2=5 if multihead else 0,
                                    pad2=2 if multihead else 0, softmax=softmax, split=projected_x3)
        else:
            self.parallel([(lambda k=k: level(k)) for k in range(len(feats))])
        conf2d = conf.view(B * P, num_classes)
        if softmax and not fused and not projected_x3:
            ops.softmax_rows(conf2d, out=conf2d)
        return loc, conf2d

    def fused_head_weight(self, loc_name, conf_name, k):
        """loc||conf rows concatenated, K-major bf16 [N_pad, kh*kw*Cin] for the fused tcgen05 head."""
        key = 'fused.%s.%s.%d' % (loc_name, conf_name, k)
        w = self.pk.get(key)
        if w is None:
            wl = self.sd['%s.%d.weight' % (loc_name, k)].detach()
            wc = self.sd['%s.%d.weight' % (conf_name, k)].detach()
            w = ops.pack_deform_head_weight(torch.cat([wl, wc], 0), self.device)   # [N = 12 + 3C, Cin, kh, kw]
            self.pk[key] = w
        return w

    def projected_head_weight(self, loc_name, conf_name, k, multihead, x3=False):
        """Per-tap projection weights (1x1 PackedConv, Cout = taps * n_pad) for tdrn_deform_head_sample."""
        key = 'proj.%s.%s.%d.%d.%d' % (loc_name, conf_name, k, int(multihead), int(x3))
        w = self.pk.get(key)
        if w is None:
            cat = lambda ln, cn: torch.cat([self.sd['%s.%d.weight' % (ln, k)].detach(),
                                            self.sd['%s.%d.weight' % (cn, k)].detach()], 0)
            w = ops.pack_deform_proj_weight(cat(loc_name, conf_name),
                                            cat(loc_name + '_2', conf_name + '_2') if multihead else None, self.device, x3=x3)
            self.pk[key] = w
        return w

    def plain_heads(self, feats, P, lv_off, num_classes, multihead, loc_name, conf_name, softmax=True,
                    keep_loc_maps=False):
        """Plain-conv loc/conf heads (RefineDet ODM refinedet_vgg.py:185-194, SSD4Scale static :111-117)."""
        B = feats[0].shape[0]
        loc = torch.empty(B, P, 4, dtype=torch.float32, device=self.device)
        conf = torch.empty(B, P, num_classes, dtype=torch.float32, device=self.device)

        def level(k):
            f = feats[k]
            self.head_into('%s.%d' % (loc_name, k), f, loc, 4, lv_off[k], P, 1)
            self.head_into('%s.%d' % (conf_name, k), f, conf, num_classes, lv_off[k], P, 1)
            if multihead:
                self.head_into('%s_2.%d' % (loc_name, k), f, loc, 4, lv_off[k], P, 2, True)
                self.head_into('%s_2.%d' % (conf_name, k), f, conf, num_classes, lv_off[k], P, 2, True)

        self.parallel([(lambda k=k: level(k)) for k in range(len(feats))])
        conf2d = conf.view(B * P, num_classes)
        if softmax:
            ops.softmax_rows(conf2d, out=conf2d)
        return loc, conf2d


def _tensors(obj):
    if torch.is_tensor(obj):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            for t in _tensors(o):
                yield t


def prior_layout(feats):
    """Prior counts/offsets for NHWC-flattened heads: index = off_k + (y*W+x)*3 + a."""
    offs, p = [], 0
    for f in feats:
        offs.append(p)
        p += f.shape[1] * f.shape[2] * NUM_BOX
    return p, offs
