"""Frame / clip sharding across ranks (one process per GPU) and the end-of-step gather.

Frames of a batch (and whole clips for TDRN: a clip's frames share the key-frame state
``(static_out[0], offset_list)``, evaluate_trn.py:434-462, so a clip never spans ranks) are independent
end to end (per-image loop layers/functions/detection.py:42).  The batch is therefore split into
contiguous per-rank slices with no collective inside the compute path; the only exchange is one
all_gather of the fixed-size ``[B_local, C, top_k, 5]`` detection buffers after Detect -- the analogue
of the implicit gather of the reference's only parallel construct, nn.DataParallel (train.py:157-159).

Backend: NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests (tests/test_shard.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous slice [lo, hi) of ``n_units`` frames (or clips) owned by ``rank``; the first
    ``n_units % world`` ranks get one extra unit, every unit is owned by exactly one rank."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError('bad rank/world %r/%r' % (rank, world))
    if n_units < 0:
        raise ValueError('n_units must be >= 0')
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank, world):
    """Rank-local slice of a batch-first tensor."""
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def shard_clips(clip_lengths, rank, world):
    """TDRN: assign whole clips to ranks -> (clip indices, frame ranges [lo, hi) in the flat frame list)."""
    lo, hi = shard_range(len(clip_lengths), rank, world)
    starts = [0]
    for n in clip_lengths:
        starts.append(starts[-1] + int(n))
    return list(range(lo, hi)), [(starts[i], starts[i + 1]) for i in range(lo, hi)]


def gather_detections(local, n_units=None, group=None, out=None):
    """All-gather the per-rank detection buffers ``[B_local, C, top_k, 5]`` -> ``[B_total, C, top_k, 5]`` in
    rank (= frame) order.  Equal shards use one all_gather_into_tensor (a single NCCL kernel on the
    compute stream, graph-capturable); ragged shards (``n_units`` not divisible by the world size) are
    padded to the largest shard and trimmed after the gather."""
    if not dist.is_available() or not dist.is_initialized():
        return local if out is None else out.copy_(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_units is None:
        n_units = local.shape[0] * world
    sizes = [shard_range(n_units, r, world)[1] - shard_range(n_units, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError('rank %d holds %d units, expected %d of %d' % (rank, local.shape[0], sizes[rank], n_units))
    tail = tuple(local.shape[1:])
    if min(sizes) == max(sizes):
        if out is None:
            out = torch.empty((n_units,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m,) + tail, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    buf = torch.empty((world * m,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * m:r * m + sizes[r]] for r in range(world)]
    res = torch.cat(parts, 0)
    return res if out is None else out.copy_(res)


class AsyncGather(object):
    """End-of-step gather that never makes a compute stream wait for another rank.

    One slot per in-flight step (a CUDA-graph instance, or simply a step index modulo ``n_slots``): ``issue(q, local)`` starts
    an asynchronous all_gather of slot q's detection buffer into its own output tensor and returns at once; ``wait(q)`` --
    called before slot q's buffers are written again, ``n_slots`` steps later -- makes the CURRENT stream wait for that
    gather.  On CUDA the gather runs behind an event recorded on the producing stream (it starts when the step's kernels
    are done, on NCCL's own stream); with gloo on the CPU (tests/test_shard.py) the same calls are plain async work handles.
    A slow rank therefore delays only gathers, up to ``n_slots - 1`` steps, never the other ranks' kernels."""

    def __init__(self, n_slots, local_shape, dtype=torch.float32, device=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        shape = (self.world * local_shape[0],) + tuple(local_shape[1:])
        self.out = [torch.empty(shape, dtype=dtype, device=device) for _ in range(n_slots)]
        self.work = [None] * n_slots
        self.cuda = device is not None and torch.device(device).type == 'cuda'
        self.stream = torch.cuda.Stream(device) if self.cuda and self.world > 1 else None
        self.done = [torch.cuda.Event() for _ in range(n_slots)] if self.stream is not None else None

    def issue(self, q, local, producer_stream=None):
        if self.world == 1:
            self.out[q].copy_(local)
            return
        if self.stream is not None:
            st = producer_stream if producer_stream is not None else torch.cuda.current_stream()
            self.done[q].record(st)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.done[q])
                self.work[q] = dist.all_gather_into_tensor(self.out[q], local, group=self.group, async_op=True)
        else:
            self.work[q] = dist.all_gather_into_tensor(self.out[q], local.contiguous(), group=self.group, async_op=True)

    def wait(self, q):
        """The current stream (CUDA) / the caller (CPU) waits for the gather that last read slot q's buffer."""
        if self.work[q] is not None:
            self.work[q].wait()
            self.work[q] = None
        return self.out[q]

    def wait_all(self):
        for q in range(len(self.work)):
            self.wait(q)
