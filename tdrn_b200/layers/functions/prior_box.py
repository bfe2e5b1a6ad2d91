"""PriorBox mirror (layers/functions/prior_box.py:5-64) over the C ABI (tdrn_prior_box, host)."""
from ... import ops


class PriorBox(object):
    def __init__(self, cfg):
        super(PriorBox, self).__init__()
        self.image_size = cfg['min_dim']
        self.num_priors = len(cfg['aspect_ratios'])
        self.variance = cfg['variance'] or [0.1]
        self.feature_maps = cfg['feature_maps']
        self.min_sizes = cfg['min_sizes']
        self.max_sizes = cfg['max_sizes']
        self.steps = cfg['steps']
        self.aspect_ratios = cfg['aspect_ratios']
        self.clip = cfg['clip']
        self.flip = cfg['flip']
        self.version = cfg['name']
        for v in self.variance:
            if v <= 0:
                raise ValueError('Variances must be greater than 0')
        self._cfg = dict(min_dim=self.image_size, feature_maps=self.feature_maps, min_sizes=self.min_sizes,
                         max_sizes=self.max_sizes, steps=self.steps, aspect_ratios=self.aspect_ratios,
                         clip=self.clip, flip=self.flip)

    def forward(self):
        """-> CPU fp32 tensor [P,4] (cx, cy, w, h), like the reference."""
        return ops.prior_box(self._cfg)
