from .l2norm import L2Norm
