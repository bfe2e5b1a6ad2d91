"""Common nn.Module behaviour of the detector mirrors (state-dict surface, precision, executor cache)."""
import os

import torch
import torch.nn as nn

from ._engine import Engine


class DetectorBase(nn.Module):
    """Holds reference-named parameters; ``forward`` runs the C-ABI executor.

    ``net.precision`` ('bf16' default, or 'fp32'; env TDRN_PRECISION overrides the default) selects
    the tcgen05 bf16 path or the fp32-accurate SIMT path.  Only phase == 'test' is implemented.
    """

    def __init__(self):
        super(DetectorBase, self).__init__()
        self.precision = os.environ.get('TDRN_PRECISION', 'bf16')
        self._engine = None

    def set_precision(self, precision):
        self.precision = precision
        self._engine = None
        return self

    def _invalidate(self):
        self._engine = None

    def load_state_dict(self, *a, **k):
        r = super(DetectorBase, self).load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super(DetectorBase, self)._apply(fn, *a, **k)
        self._invalidate()
        return r

    def load_weights(self, base_file):
        other, ext = os.path.splitext(base_file)
        if ext in ('.pkl', '.pth'):
            print('Loading weights into state dict...')
            self.load_state_dict(torch.load(base_file, map_location=lambda storage, loc: storage))
            print('Finished!')
        else:
            print('Sorry only .pth and .pkl files supported.')

    def engine(self):
        if self.phase != 'test':
            raise NotImplementedError("tdrn_b200 implements inference only (phase='test')")
        if self._engine is None or self._engine.precision != self.precision:
            self._engine = Engine(self, self.precision)
        return self._engine

    @staticmethod
    def _check_input(x):
        if not torch.is_tensor(x) or x.dim() != 4 or x.size(1) != 3:
            raise ValueError('expected input of shape [B,3,S,S]')
        if not x.is_cuda:
            raise NotImplementedError('input must be a CUDA tensor: tdrn_b200 has no CPU path')
        return x.float().contiguous()
