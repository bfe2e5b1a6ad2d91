#!/usr/bin/env python
"""bench.py -- frames/s of the DualRefineDet-VGGBN-320 inference hot path (net(x) + Detect) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1: one rank per GPU)

A "step" is one pass of the hot path over one batch of 32 synthetic frames per GPU (BASELINE.json
configs[1]): conv backbone + ARM/ODM heads (deformable, multihead as scripts/batch_eval.sh ships it),
softmax, two-stage decode, per-class NMS, top-k.  Frames shard by batch across ranks (weak scaling, no
collective inside the compute path; one NCCL all_gather of the fixed-size detection buffers per step).

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM (CUDA-graph replay);
`e2e` = the same through the public API with HOST buffers (pinned H2D of the frames + D2H of the
detections inside the timed region); `roofline` = the tcgen05 implicit-GEMM conv kernel (dominant),
algorithmic FLOPs / CUDA-event time; `cpu_baseline` = the oracle port of the reference's CPU path timed
on this box's host cores.  `--impl reference` times only that CPU path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'frames/sec DualRefineDet-VGGBN-320 b32'
UNIT = 'frames/s'
BATCH = 32
SIZE = 320
NUM_CLASSES = 21
MODEL_KW = dict(num_classes=NUM_CLASSES, def_groups=1, bn=True, multihead=True)
DETECT_KW = dict(top_k=200, conf_thresh=0.01, nms_thresh=0.45)      # scripts/batch_eval.sh:2-4
GFLOP_PER_FRAME = 77.466                                           # SURVEY.md 8d (multihead, conv + deform)


# diagnosis only (never set by the driver): TDRN_BENCH_E2E_SKIP=h2d,d2h drops those copies from the e2e leg to attribute its
# gap to the device leg; a run with it set is not a valid e2e number and says so in its JSON line
_DIAG_SKIP = [t for t in os.environ.get('TDRN_BENCH_E2E_SKIP', '').split(',') if t]


def config_dict(n_gpus, inflight=2):
    return {'workload': 'DualRefineDet-VGGBN 320x320 VOC-21 batch %d per GPU, multihead deformable ODM, '
                        'net(x)+Detect(top_k 200, conf 0.01, nms 0.45)' % BATCH,
            'global_batch': BATCH * n_gpus, 'per_gpu_batch': BATCH, 'input': '[B,3,320,320] fp32 N(0,1)',
            'weights': 'seeded random init, randomised BN statistics (tdrn_b200.utils.synthetic.randomize_ seed 0)',
            'parallelism': 'dp%d (frames sharded by batch, weights replicated)' % n_gpus,
            'l2': 'rotating 4 distinct input batches (157 MB) and >1 GB of per-step activations exceed the 126 MB L2',
            'pipelining': 'CUDA-graph replay (one instance per resident input batch), %d step(s) in flight (with 2, step i+1 trunk '
                          'overlaps the latency-bound tail of step i)' % inflight}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'bf16_tflops': p['bf16_tflops'], 'bf16_tflops_sustained': p['bf16_tflops_sustained'],
                'hbm_gbs': p['hbm_gbs'], 'source': 'measured (MEASURED_PEAKS.json)'}
    return {'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


def conv_traffic():
    """DRAM bytes per launch of the conv family from the committed ncu --set full capture (profiles/conv_traffic.json,
    written by scripts/ncu_traffic.py from the .ncu-rep of the same bench command); None when no capture is committed."""
    path = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path))['dram_bytes_per_launch']
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's own CPU implementation)
# ------------------------------------------------------------------------------------------------------
class CpuReference(object):
    def __init__(self):
        import torch
        from oracle import model_ref as M, detect_ref as D
        self.torch, self.M, self.D = torch, M, D
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = {k: v.detach().cpu() for k, v in build_synthetic_net().state_dict().items()}   # same weights as the GPU arm
        self.priors = D.prior_box(D.VOC_320)

    def step(self, x):
        """x [b,3,320,320] -> detections; torch CPU convs on all cores, Detect on one core like the reference."""
        import numpy as np
        from oracle import c_oracle as C
        arm_loc, _, loc, conf = self.M.drn_vgg_forward(self.sd, x, **MODEL_KW)
        boxes = C.decode(loc.numpy(), self.priors.numpy(), arm_loc.numpy())
        return C.detect(boxes, conf.numpy(), np.array([320.] * 4, np.float32), NUM_CLASSES, DETECT_KW['top_k'],
                        DETECT_KW['conf_thresh'], DETECT_KW['nms_thresh'])

    def run(self, steps, warmup, frames_per_step=1):
        from tdrn_b200.utils.synthetic import frames
        x = frames(frames_per_step, SIZE, seed=11)
        for _ in range(warmup):
            self.step(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(x)
        dt = time.perf_counter() - t0
        return frames_per_step * steps / dt, dt / steps * 1e3


def build_synthetic_net():
    from tdrn_b200.model import dualrefinedet_vggbn as V
    from tdrn_b200.utils.synthetic import randomize_
    return randomize_(V.build_net('test', SIZE, **MODEL_KW), seed=0).eval()


def run_reference(args, rank):
    if rank != 0:
        return
    ref = CpuReference()
    fps, ms = ref.run(args.steps, args.warmup, 1)
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config_dict(args.gpus),
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': ref.cores, 'kind': 'port',
                             'sample': '1 frame per step through the oracle port (torch CPU convs on all cores + '
                                       'scalar C Detect/NMS on one core, as the reference runs it); frames/s is '
                                       'batch-independent on CPU'},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi polled every 20 ms for the whole run (its start-up takes 0.1-0.5 s, far longer than the timed
    region, so it is launched before the warm-up); begin()/end() bracket the timed legs and only samples that
    arrived inside the bracket are reported (falling back to the closest ones if the bracket was shorter than a poll)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.03]
        window = 'timed legs'
        if not rows and self.rows and self.t0 is not None:     # bracket shorter than one poll: closest samples
            mid = 0.5 * (self.t0 + self.t1)
            rows = [r for t, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            window = 'closest to the timed legs'
        sm, mx, pw, reasons = [], None, [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'power_w_max': max(pw) if pw else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from tdrn_b200 import ops, _lib
    from tdrn_b200.utils.synthetic import frames as make_frames
    from tdrn_b200.layers.functions import Detect, PriorBox
    from tdrn_b200.data import mb_cfg
    from tdrn_b200.utils.shard import gather_detections

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        # stdout carries ONE JSON line, but NCCL printf()s its version banner to stdout when the communicator is created:
        # point fd 1 at stderr while the process group (and its communicator, via the first collective) comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    net = build_synthetic_net().to(dev).set_precision(args.precision)
    det = Detect(NUM_CLASSES, 0, DETECT_KW['top_k'], DETECT_KW['conf_thresh'], DETECT_KW['nms_thresh'])
    priors = PriorBox(mb_cfg['VOC_320']).forward().to(dev)

    n_in = 4
    host_x = [make_frames(BATCH, SIZE, seed=100 + rank * n_in + i).pin_memory() for i in range(n_in)]
    dev_x = [h.to(dev) for h in host_x]
    host_out = torch.empty(BATCH, NUM_CLASSES, DETECT_KW['top_k'], 5).pin_memory()
    def hot_path(x):
        arm_loc, _, loc, conf = net(x)
        return det.forward(loc, conf, priors, arm_loc_data=arm_loc)

    # Two steps are kept in flight on two streams (a serving loop would do the same): the tail of a step (FPN chain,
    # small pyramid levels, NMS) is latency-bound and leaves SMs idle that the next step's trunk can use.  A "step" is still
    # one pass over one batch of 32 frames; K steps are timed; --inflight 1 gives the strictly serial number (~5 % slower).
    # There is one captured graph instance per resident input batch (n_in = 4): instance q reads xs[q] in place, so the
    # device leg moves no input bytes at all and the e2e leg's H2D lands directly in the graph's input (no staging copy
    # that would queue behind the 1.5 ms H2D on a copy engine).  Step i runs instance i % 4 on stream i % 2.
    n_fl = 1 if args.no_graph else max(1, args.inflight)
    # instance q is captured on stream q % n_fl and must always replay there (its Detect workspace belongs to that stream, and
    # consecutive uses of one instance must be stream-ordered): the instance count is a multiple of the stream count
    n_slots = 1 if args.no_graph else n_fl * ((max(n_in, n_fl + 2) + n_fl - 1) // n_fl)
    streams = [torch.cuda.Stream(dev) for _ in range(n_fl)]
    stream = streams[0]
    xs = [dev_x[q % n_in].clone() for q in range(n_slots)]
    graphs, static_outs = [], []
    torch.cuda.synchronize()
    with torch.cuda.stream(stream), torch.no_grad():
        for i in range(3):                      # eager warm-up: packs weights, sizes workspaces, loads kernels
            hot_path(dev_x[i % n_in])
        stream.synchronize()
        l0 = _lib.launch_count()
        hot_path(dev_x[0])
        launches_per_step = _lib.launch_count() - l0
        stream.synchronize()
    for k in range(1, n_fl):
        with torch.cuda.stream(streams[k]), torch.no_grad():
            hot_path(xs[0])                     # eager once on this stream: its own Detect workspace exists before capture
        streams[k].synchronize()
    for q in range(n_slots):
        st = streams[q % n_fl]
        with torch.cuda.stream(st), torch.no_grad():
            if args.no_graph:
                graphs.append(None)
                static_outs.append(hot_path(xs[q]))
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    static_outs.append(hot_path(xs[q]))
                graphs.append(g)
        st.synchronize()
    gathered = [torch.empty(world * BATCH, NUM_CLASSES, DETECT_KW['top_k'], 5, device=dev) for _ in range(n_slots)] if world > 1 else None

    def replay(q):
        if args.no_graph:
            with torch.no_grad():
                static_outs[q].copy_(hot_path(xs[q]))
        else:
            graphs[q].replay()

    def step_device(i):
        q = i % n_slots
        with torch.cuda.stream(streams[i % n_fl]):
            if n_slots < n_in:
                xs[q].copy_(dev_x[i % n_in], non_blocking=True)         # eager profiling mode: one input buffer
            replay(q)                                                   # the frames are already in HBM (xs[q])
            if world > 1:
                gather_detections(static_outs[q], out=gathered[q])

    # e2e: software-pipelined like a production feeder -- the pinned-host -> device copy of step i + n_fl runs on a copy
    # stream while steps i, i + 1 compute; every step still pays its own H2D (39 MB) and D2H (2.7 MB) inside the timed
    # region, they just overlap with the neighbouring steps' kernels instead of serialising with them.
    copy_stream = torch.cuda.Stream(dev)                      # host -> device feeder
    d2h_stream = torch.cuda.Stream(dev)                       # detections -> host (own stream: a D2H waiting for step i
                                                              # must not hold back the H2D of step i+2)
    in_ready = [torch.cuda.Event() for _ in range(n_slots)]   # H2D into xs[q] finished
    done_ev = [torch.cuda.Event() for _ in range(n_slots)]    # the step that read xs[q] / wrote static_outs[q] has finished
    out_copied = [torch.cuda.Event() for _ in range(n_slots)]
    pipe = {'primed': -1}

    def prefetch(i):
        q = i % n_slots
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done_ev[q])                # the previous user of this instance's input is done
            if 'h2d' not in _DIAG_SKIP:
                xs[q].copy_(host_x[i % n_in], non_blocking=True)
            in_ready[q].record(copy_stream)
        pipe['primed'] = i

    def step_e2e(i):
        depth = min(n_fl, n_slots - 1)
        for j in range(pipe['primed'] + 1, i + 1):
            prefetch(j)
        q = i % n_slots
        st = streams[i % n_fl]
        with torch.cuda.stream(st):
            st.wait_event(in_ready[q])
            st.wait_event(out_copied[q])                            # this instance's previous detections have left the device
            replay(q)
            if world > 1:
                gather_detections(static_outs[q], out=gathered[q])
            done_ev[q].record(st)
        for j in range(pipe['primed'] + 1, i + depth + 1):          # the H2D of the next steps overlaps the kernels in flight
            prefetch(j)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done_ev[q])
            if 'd2h' not in _DIAG_SKIP:
                host_out.copy_(static_outs[q], non_blocking=True)   # detections -> pinned host
            out_copied[q].record(d2h_stream)

    def timed(step_fn, steps, warmup):
        pipe['primed'] = -1
        w = max(warmup, 12)                                # W is a minimum: a few more replays let the clocks settle
        for i in range(w):
            step_fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for st in streams[1:]:
            st.wait_event(e0)
        copy_stream.wait_event(e0)
        d2h_stream.wait_event(e0)
        for i in range(steps):
            step_fn(w + i)
        for st in streams[1:]:
            stream.wait_stream(st)
        stream.wait_stream(copy_stream)
        stream.wait_stream(d2h_stream)                      # e2e: the last step's D2H is inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler.begin()
    ms_dev = timed(step_device, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None          # sampled every 20 ms across both timed legs

    # ---- overlapped replay must give what a serial eager pass gives (bit for bit: every kernel is deterministic) ----
    inflight_ok = None
    if not args.no_graph:
        for i in range(2 * n_slots):
            step_device(i)                                  # instance q ends up holding the detections of xs[q]
        torch.cuda.synchronize()
        with torch.cuda.stream(stream), torch.no_grad():
            inflight_ok = all(bool(torch.equal(static_outs[q], hot_path(xs[q]))) for q in range(n_slots))
        torch.cuda.synchronize()
        if not inflight_ok:       # reported in the JSON line (inflight_replay_matches_serial: false), never silently dropped
            sys.stderr.write('bench: WARNING detections of overlapped graph replays differ from a serial eager pass\n')

    # ---- roofline leg: per-call CUDA events on the launching stream, eager (same kernels as the graph) ----
    roof = None
    if rank == 0:
        # per-kernel timing needs the kernels serialised on one stream: switch the engine's fork/join branches
        # off for this leg only (the timed legs above run the multi-stream graph)
        net.engine().multi_stream = False
        with torch.cuda.stream(stream), torch.no_grad():
            hot_path(dev_x[1])
            stream.synchronize()
            ops.prof_begin()
            for i in range(max(2, min(args.steps, 5))):
                hot_path(dev_x[i % n_in])
            rec = ops.prof_end()
        net.engine().multi_stream = True
        agg, detail = {}, {}
        for label, work, ms in rec:
            a = agg.setdefault(label.split('|')[0], [0.0, 0.0, 0])
            a[0] += work; a[1] += ms; a[2] += 1
            dd = detail.setdefault(label, [0.0, 0.0, 0])
            dd[0] += work; dd[1] += ms; dd[2] += 1
        if args.detail:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1][1]):
                sys.stderr.write('%-48s n=%3d  %8.4f ms/launch  %10.2f G(work)/s\n' % (k, v[2], v[1] / v[2], v[0] / (v[1] * 1e-3) / 1e9))
        pk = peaks()
        tot_ms = sum(a[1] for a in agg.values())
        tc = agg.get('conv_tc')
        if tc:
            tflops = tc[0] / (tc[1] * 1e-3) / 1e12
            roof = {'kernel': 'tcgen05 implicit-GEMM conv family (conv_stem_pair_kernel = conv1_1+conv1_2 fused, conv_halo_kernel, '
                              'conv_halo_stream_kernel, conv_tc_kernel): all conv launches of the step, algorithmic FLOPs / summed '
                              'CUDA-event time', 'bound': 'tensor',
                    'achieved': tflops, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': tflops / pk['bf16_tflops_sustained'], 'traffic': conv_traffic(),
                    'peak_source': pk['source'] + ', sustained bf16 (kernel timed inside a long step)',
                    'launches_per_step': tc[2] // max(2, min(args.steps, 5)),
                    'avg_launch_ms': tc[1] / tc[2], 'share_of_step': tc[1] / tot_ms}
        breakdown = {k: {'ms_per_step': v[1] / max(2, min(args.steps, 5)), 'launches': v[2] // max(2, min(args.steps, 5)),
                         'work_per_step': v[0] / max(2, min(args.steps, 5))} for k, v in agg.items()}
        if 'detect' in agg:
            d = agg['detect']
            breakdown['detect']['achieved_gbs'] = d[0] / (d[1] * 1e-3) / 1e9
            breakdown['detect']['frac_of_hbm'] = breakdown['detect']['achieved_gbs'] / pk['hbm_gbs']
            if not args.no_graph:          # (the ncu launch-list runs use --no-graph: keep their tail = one step)
                # regime T: trained-like scores -- background logit +7.7, so ~1.5 % of the (prior, class) scores pass conf_thresh --
                # where Detect is bound by HBM traffic rather than by the IoU scan of 6 375 candidates per class (random init)
                g = torch.Generator(device='cpu').manual_seed(7)
                logits = torch.randn(BATCH * priors.shape[0], NUM_CLASSES, generator=g)
                logits[:, 0] += 7.7
                conf_t = torch.softmax(logits, 1).to(dev)
                loc_t = torch.randn(BATCH, priors.shape[0], 4, generator=g).to(dev)
                arm_t = (0.5 * torch.randn(BATCH, priors.shape[0], 4, generator=g)).to(dev)
                with torch.cuda.stream(stream), torch.no_grad():
                    for _ in range(3):
                        det.forward(loc_t, conf_t, priors, arm_loc_data=arm_t)
                    ops.prof_begin()
                    for _ in range(10):
                        det.forward(loc_t, conf_t, priors, arm_loc_data=arm_t)
                    rec_t = ops.prof_end()
                ms_t = sum(r[2] for r in rec_t) / len(rec_t)
                breakdown['detect']['trained_like'] = {
                    'ms_per_step': ms_t, 'candidates_frac': float((conf_t[:, 1:] > DETECT_KW['conf_thresh']).float().mean().item()),
                    'achieved_gbs': rec_t[0][1] / (ms_t * 1e-3) / 1e9, 'frac_of_hbm': rec_t[0][1] / (ms_t * 1e-3) / 1e9 / pk['hbm_gbs']}
        if 'deform_head_tc' in agg:
            d = agg['deform_head_tc']
            breakdown['deform_head_tc']['achieved_tflops'] = d[0] / (d[1] * 1e-3) / 1e12
            breakdown['deform_head_tc']['frac_of_tensor_peak'] = d[0] / (d[1] * 1e-3) / 1e12 / pk['bf16_tflops_sustained']
            # the projection GEMM of pyramid level 0 writes B*1600*34*80 bf16 projections: HBM-write bound
            pj = [v for k, v in detail.items() if k.startswith('deform_head_tc') and '@40x40' in k and k.endswith('project')]
            if pj:
                ms_pj = pj[0][1] / pj[0][2]
                wr = BATCH * 1600 * 34 * 80 * 2
                breakdown['deform_head_tc']['projection_level0'] = {
                    'ms': ms_pj, 'bytes_written': wr, 'achieved_write_gbs': wr / (ms_pj * 1e-3) / 1e9,
                    'note': 'pure 278 MB memset on this B200: 3750 GB/s (profiles/probe_gemm_bound.txt); cuBLAS on the same GEMM 0.093 ms'}
    if world > 1:
        dist.barrier()

    if rank == 0:
        frames = BATCH * world * args.steps
        value = frames / (ms_dev * 1e-3)
        e2e_v = frames / (ms_e2e * 1e-3)
        cpu = None
        if world == 1 and not args.no_cpu:
            ref = CpuReference()
            fps, _ = ref.run(steps=6, warmup=1, frames_per_step=2)
            cpu = {'value': fps, 'unit': UNIT, 'cores': ref.cores, 'kind': 'port',
                   'sample': '6 steps x 2 frames of the same workload through the oracle port (torch CPU convs on all '
                             'cores + scalar C Detect/NMS), after 1 warm-up step'}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic', 'config': config_dict(world, n_fl),
                'e2e': {'value': e2e_v, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                        'h2d_bytes_per_step': int(xs[0].numel() * 4), 'd2h_bytes_per_step': int(host_out.numel() * 4)},
                'gpu_launches': int(launches_per_step * args.steps),
                'launches_per_step': int(launches_per_step),
                'inflight_replay_matches_serial': inflight_ok,
                **({'INVALID_e2e_diagnosis_skip': _DIAG_SKIP} if _DIAG_SKIP else {}),
                'tflops_per_gpu_whole_step': GFLOP_PER_FRAME * BATCH / (ms_dev / args.steps),   # GFLOP / ms == TFLOP/s
                'roofline': roof, 'kernel_breakdown': breakdown, 'cpu_baseline': cpu, 'clocks': clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='tdrn_b200', choices=['tdrn_b200', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--inflight', type=int, default=2, help='steps kept in flight on separate streams / graph instances (1 = strictly serial)')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay (profiling)')
    ap.add_argument('--detail', action='store_true', help='print the per-layer CUDA-event breakdown to stderr')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-exec ourselves one rank per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
