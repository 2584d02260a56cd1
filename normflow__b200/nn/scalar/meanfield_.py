"""MeanFieldNet_: a distribution convertor acting on the lattice average of a field
(reference src/nn/scalar/meanfield_.py)."""

import numpy as np

from .modules_ import DistConvertor_
from .._core import Module_
from ... import _ops


class MeanFieldNet_(Module_):
    """Transforms sqrt(V) * mean(x) with a DistConvertor_ and puts the result back
    (meanfield_.py:18-48).  With `rvol` given, x already IS the per-sample mean
    (shape [B, 1, ...]) and rvol = sqrt(V) -- the way PSDBlock_ calls it."""

    def __init__(self, dc_, label='mean-field'):
        super().__init__(label)
        self.dc_ = dc_

    def _convert(self, x, log0, rvol, inverse):
        if rvol is not None:
            fn = self.dc_.backward if inverse else self.dc_.forward
            y, log0 = fn(x * rvol, log0)
            return y / rvol, log0
        rvol = float(np.prod(x.shape[1:])) ** 0.5
        x_mean = _ops.sample_mean(x)
        fn = self.dc_.backward if inverse else self.dc_.forward
        new_scaled, log0 = fn((x_mean * rvol).reshape(-1, 1), log0)
        return _ops.sample_shift(x, new_scaled.reshape(-1) / rvol - x_mean), log0

    def forward(self, x, log0=0, rvol=None):
        return self._convert(x, log0, rvol, inverse=False)

    def backward(self, x, log0=0, rvol=None):
        return self._convert(x, log0, rvol, inverse=True)

    def _hack(self, x, log0=0):
        """(mean, log) before and after the convertor (meanfield_.py:50-61)."""
        rvol = float(np.prod(x.shape[1:])) ** 0.5
        x_mean = _ops.sample_mean(x)
        stack = [(x_mean, log0)]
        new_scaled, log0 = self.dc_.forward((x_mean * rvol).reshape(-1, 1), log0)
        stack.append((new_scaled.reshape(-1) / rvol, log0))
        return stack

    @staticmethod
    def build(knots_len=10, **kwargs):
        return MeanFieldNet_(DistConvertor_(knots_len, **kwargs))
