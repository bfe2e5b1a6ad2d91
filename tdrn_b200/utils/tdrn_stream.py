"""TDRN streaming executor: the per-frame key-frame loop of the reference's video drivers
(evaluate_trn.py:434-467, test_video_trn.py:81-103) around the static and temporal SSD4Scale detectors.

On a key frame (first frame of a video, or every ``interval`` frames) the static detector runs; its (loosened)
regression becomes the ARM stage of Detect for the following frames and its raw loc maps drive the temporal net's
offset convs.  The temporal net returns its offsets on that frame; they are cached and re-used (``offset_list``)
until the next key frame.  State per stream: ``(static_out[0], ref_loc, offset_list, current_i, video name)``.
"""


class TDRNStream(object):
    def __init__(self, static_net, net, detector, priors, interval=4, loose=1.0, deform=True):
        self.static_net, self.net, self.detector, self.priors = static_net, net, detector, priors
        self.interval, self.loose, self.deform = int(interval), float(loose), bool(deform)
        self.reset()

    def reset(self):
        self.pre_video_name = None
        self.current_i = 0
        self.offset_list = list()
        self.ref_loc = list()
        self.static_out = None

    def is_key_frame(self, video_name=None):
        return self.static_out is None or video_name != self.pre_video_name or self.current_i % self.interval == 0

    def step(self, x, video_name=None):
        """x [1,3,S,S] (or a batch of frames that share the key-frame state) -> Detect output [B,C,top_k,5]."""
        if self.is_key_frame(video_name):                                            # evaluate_trn.py:450
            self.static_out = list(self.static_net(x, ret_loc=self.deform))          # :451
            self.static_out[0] = self.static_out[0] * self.loose                     # :452
            if self.deform:
                self.ref_loc = self.static_out[2]                                    # :454
                self.offset_list = list()                                            # :455
            if video_name != self.pre_video_name:                                    # :456-458
                self.pre_video_name = video_name
                self.current_i = 0
        out = self.net(x, ref_loc=self.ref_loc, offset_list=self.offset_list,
                       ret_off=bool(self.deform and not self.offset_list))           # :459
        if len(out) == 3:                                                            # :460-462
            self.offset_list = out[2]
            self.ref_loc = list()
        detections = self.detector.forward(out[0], out[1], self.priors, arm_loc_data=self.static_out[0])   # :465
        self.current_i += 1                                                          # :467
        return detections



class GraphedTDRNStream(TDRNStream):
    """The same per-frame loop with the two kinds of frame captured as CUDA graphs (batch-1 video latency: a frame is ~130 kernel
    launches on a key frame, ~70 otherwise; launched one by one the host would be the bottleneck).

      key frame     : static net (side stream)  ||  temporal trunk (main stream)  ->  offsets from the static regression ->
                      temporal deformable heads -> Detect against the (loosened) static regression
      other frames  : temporal net with the key frame's cached offsets -> Detect against the cached regression

    The caches (`static_out[0]`, `offset_list`) are tensors owned by the key-frame graph; the other graph reads them in place.
    ``step(frame, video_name)`` copies the frame into the graphs' input and replays one of them; the returned tensor is the
    graph's output buffer (consume it before the next step).  `frame` is ``[1,3,S,S]`` fp32, or -- with ``mean`` given -- a
    uint8 ``[1,H,W,3]`` camera frame that `tdrn_preprocess` (base_transform) turns into the network input inside the graph."""

    def __init__(self, static_net, net, detector, priors, size=320, interval=4, loose=1.0, mean=None, frame_hw=None):
        import torch
        from .. import ops
        super(GraphedTDRNStream, self).__init__(static_net, net, detector, priors, interval=interval, loose=loose, deform=True)
        dev = priors.device
        self._x = torch.zeros(1, 3, size, size, device=dev)
        self._u8 = None
        if mean is not None:
            h, w = frame_hw if frame_hw is not None else (size, size)
            self._u8 = torch.zeros(1, h, w, 3, dtype=torch.uint8, device=dev)
        self._stream = torch.cuda.Stream(dev)
        side = torch.cuda.Stream(dev)

        def prep():
            if self._u8 is not None:
                ops.preprocess(self._u8, size, mean, out=self._x)

        def key_frame():
            prep()
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event(); fork.record(main)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                s_out = list(self.static_net(self._x, ret_loc=True))                 # evaluate_trn.py:451
                arm = s_out[0] * self.loose                                          # :452
                done = torch.cuda.Event(); done.record(side)
            src = self.net.trunk(self._x)                                            # needs nothing from the static net yet
            main.wait_event(done)
            for t in [arm] + list(s_out[2]):
                t.record_stream(main)
            out = self.net(self._x, ref_loc=s_out[2], offset_list=[], ret_off=True, _sources=src)     # :459-462
            det = self.detector.forward(out[0], out[1], self.priors, arm_loc_data=arm)           # :465
            return arm, out[2], det

        def other_frame(arm, offs):
            prep()
            out = self.net(self._x, ref_loc=[], offset_list=offs)
            return self.detector.forward(out[0], out[1], self.priors, arm_loc_data=arm)

        with torch.cuda.stream(self._stream), torch.no_grad():
            arm, offs, _ = key_frame()                      # eager warm-up: packs weights, sizes the Detect workspace of this stream
            other_frame(arm, offs)
            self._stream.synchronize()
            self._g_key = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_key, stream=self._stream):
                self._arm, self._offs, self._det_key = key_frame()
            self._g_other = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_other, stream=self._stream):
                self._det_other = other_frame(self._arm, self._offs)
        self._stream.synchronize()

    def step(self, frame, video_name=None):
        import torch
        key = self.is_key_frame(video_name)
        with torch.cuda.stream(self._stream):
            (self._u8 if frame.dtype == torch.uint8 else self._x).copy_(frame, non_blocking=True)
            if key:
                self._g_key.replay()
                self.static_out = [self._arm]
                self.offset_list = self._offs
                if video_name != self.pre_video_name:
                    self.pre_video_name = video_name
                    self.current_i = 0
            else:
                self._g_other.replay()
        self.current_i += 1
        return self._det_key if key else self._det_other

    def synchronize(self):
        self._stream.synchronize()
