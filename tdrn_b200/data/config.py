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
