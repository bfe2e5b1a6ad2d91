"""Prior-box dictionaries the hot path needs (values of the reference's data/config.py:57-81, :260-261)."""

VOC_320 = {
    'feature_maps': [40, 20, 10, 5], 'min_dim': 320, 'steps': [8, 16, 32, 64],
    'min_sizes': [32, 64, 128, 256], 'max_sizes': [], 'aspect_ratios': [[2], [2], [2], [2]],
    'variance': [0.1, 0.2], 'clip': True, 'flip': True, 'name': 'VOC_320',
}

VOC_512_RefineDet = {
    'feature_maps': [64, 32, 16, 8], 'min_dim': 512, 'steps': [8, 16, 32, 64],
    'min_sizes': [32, 64, 128, 256], 'max_sizes': [], 'aspect_ratios': [[2], [2], [2], [2]],
    'variance': [0.1, 0.2], 'clip': True, 'flip': True, 'name': 'VOC_512_RefineDet',
}


def _ssd_prior_cfg(name, size, min_sizes, max_sizes, aspect_ratios, flip=True):
    """The SSD / RFB-style dictionaries of data/config.py:5-55,83-137 (six levels at 300, seven at 512, with max_sizes)."""
    n = len(min_sizes)
    steps = [8, 16, 32, 64, 100, 300] if size == 300 else [8, 16, 32, 64, 128, 256, 512]
    fmaps = [38, 19, 10, 5, 3, 1] if size == 300 else [64, 32, 16, 8, 4, 2, 1]
    assert n == len(steps) == len(max_sizes) == len(aspect_ratios)
    return {'feature_maps': fmaps, 'min_dim': size, 'steps': steps, 'min_sizes': min_sizes, 'max_sizes': max_sizes,
            'aspect_ratios': aspect_ratios, 'variance': [0.1, 0.2], 'clip': True, 'flip': flip, 'name': name}


_S300 = ([30, 60, 111, 162, 213, 264], [60, 111, 162, 213, 264, 315])
VOC_300 = _ssd_prior_cfg('VOC_300', 300, *_S300, [[2], [2, 3], [2, 3], [2, 3], [2], [2]])
VOC_300_RFB = _ssd_prior_cfg('VOC_300_RFB', 300, *_S300, [[2, 3], [2, 3], [2, 3], [2, 3], [2], [2]])
MOT_300 = _ssd_prior_cfg('MOT_300', 300, *_S300, [[1 / 2, 1 / 3, 1 / 4]] * 6, flip=False)
COCO_300 = _ssd_prior_cfg('COCO_300', 300, [21, 45, 99, 153, 207, 261], [45, 99, 153, 207, 261, 315],
                          [[2, 3], [2, 3], [2, 3], [2, 3], [2], [2]])
_AR512 = [[2, 3], [2, 3], [2, 3], [2, 3], [2, 3], [2], [2]]
VOC_512 = _ssd_prior_cfg('VOC_512', 512, [35.84, 76.8, 153.6, 230.4, 307.2, 384.0, 460.8],
                         [76.8, 153.6, 230.4, 307.2, 384.0, 460.8, 537.6], _AR512)
COCO_512 = _ssd_prior_cfg('COCO_512', 512, [20.48, 51.2, 133.12, 215.04, 296.96, 378.88, 460.8],
                          [51.2, 133.12, 215.04, 296.96, 378.88, 460.8, 542.72], _AR512)

# data/config.py:257-258: every dictionary the reference's `mb_cfg` holds (the hot-path models use VOC_320 and
# VOC_512_RefineDet; the SSD / RFB entries are here so that `from data import mb_cfg` keeps working unchanged)
mb_cfg = {'VOC_300': VOC_300, 'VOC_300_RFB': VOC_300_RFB, 'VOC_320': VOC_320, 'VOC_512': VOC_512, 'MOT_300': MOT_300,
          'COCO_300': COCO_300, 'COCO_512': COCO_512, 'VOC_512_RefineDet': VOC_512_RefineDet}


def _refinedet_prior_cfg(size, name):
    """The multi-scale prior dictionaries of data/config.py:139-258 differ only in min_dim and feature_maps = size / steps."""
    return {'feature_maps': [size // 8, size // 16, size // 32, size // 64], 'min_dim': size, 'steps': [8, 16, 32, 64],
            'min_sizes': [32, 64, 128, 256], 'max_sizes': [], 'aspect_ratios': [[2], [2], [2], [2]],
            'variance': [0.1, 0.2], 'clip': True, 'flip': True, 'name': name}


# data/config.py:260-261 and multi_eval.py:21-24
multi_cfg = {'192': _refinedet_prior_cfg(192, 'VOC_192'), '320': VOC_320, '384': _refinedet_prior_cfg(384, 'VOC_384'),
             '448': _refinedet_prior_cfg(448, 'VOC_448'), '512': _refinedet_prior_cfg(512, 'VOC_512_s'),
             '576': _refinedet_prior_cfg(576, 'VOC_576'), '704': _refinedet_prior_cfg(704, 'VOC_704')}
multi_cfg_512 = {'320': _refinedet_prior_cfg(320, 'VOC_512_RefineDet_06'), '512': VOC_512_RefineDet,
                 '640': _refinedet_prior_cfg(640, 'VOC_512_RefineDet_12'),
                 '1216': _refinedet_prior_cfg(1216, 'VOC_512_RefineDet_22')}
multi_scale = {'320': [192, 320, 384, 448, 512, 576, 704], '512': [320, 512, 640, 1216]}
