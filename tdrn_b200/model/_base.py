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
        self._engine_stamp = None

    def set_precision(self, precision):
        self.precision = precision
        self._engine = None
        return self

    def _invalidate(self):
        self._engine = None
        self.__dict__['_stamp_tensors'] = None

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

    def _param_stamp(self):
        """(data_ptr, in-place version counter) of every parameter and buffer: changes whenever a weight is rebound or
        written in place (``p.copy_``, ``net.apply(init)``, optimizer / EMA steps, manual BN-statistics edits).  Writes through
        ``p.data`` bypass PyTorch's version counters and cannot be seen: call ``net.refresh()`` after those."""
        ts = self.__dict__.get('_stamp_tensors')
        if ts is None:                               # (the module walk is the expensive part: cached until _invalidate)
            ts = self.__dict__['_stamp_tensors'] = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in ts)

    def refresh(self):
        """Drop the packed (BN-folded, bf16 / split) weights; they are rebuilt from the current parameters at the next call."""
        self._invalidate()
        return self

    def engine(self):
        if self.phase != 'test':
            raise NotImplementedError("tdrn_b200 implements inference only (phase='test')")
        # The Engine holds BN-folded, re-packed copies of the weights.  They are rebuilt when the precision changes, after
        # load_state_dict / .to(), and when any parameter or buffer was modified in place since they were packed.  (Captured
        # CUDA graphs hold the old packed buffers: re-capture after changing weights.)
        stamp = self._param_stamp()
        if self._engine is None or self._engine.precision != self.precision or self._engine_stamp != stamp:
            self._engine = Engine(self, self.precision)
            self._engine_stamp = stamp
        return self._engine

    def __deepcopy__(self, memo):
        return self._clone_without_engine(memo)

    def _clone_without_engine(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_engine' else copy.deepcopy(v, memo)
        return new

    def _replicate_for_data_parallel(self):
        # nn.DataParallel replicas are shallow copies: they must not share the device-0 executor
        replica = super(DetectorBase, self)._replicate_for_data_parallel()
        replica._engine = None
        return replica

    @staticmethod
    def _check_input(x):
        if not torch.is_tensor(x) or x.dim() != 4 or x.size(1) != 3:
            raise ValueError('expected input of shape [B,3,S,S]')
        if not x.is_cuda:
            raise NotImplementedError('input must be a CUDA tensor: tdrn_b200 has no CPU path')
        return x.float().contiguous()
