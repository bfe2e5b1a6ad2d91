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
