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

mb_cfg = {'VOC_320': VOC_320, 'VOC_512_RefineDet': VOC_512_RefineDet}


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
