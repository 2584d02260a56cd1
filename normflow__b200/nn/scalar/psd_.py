"""PSDBlock_: power-spectral-density block = mean-field net on the lattice average plus
FFTNet_ on the fluctuations (reference src/nn/scalar/psd_.py).

With an FFTNet_ the block is evaluated in the Fourier domain in one pass: the lattice
average is the zero mode of rfftn(x) (divided by V), subtracting it before the
transform is the same as dropping that mode, and adding the mean-field output y_mf
afterwards is the same as writing V * y_mf into it.  So the whole block is

    rfftn  ->  [zero mode -> mean-field net]  ->  one kernel: spectrum * w, zero mode := V y_mf  ->  irfftn

instead of mean, subtract, rfftn, multiply, irfftn, add.
"""

import numpy as np
import torch

from .fftflow_ import FFTNet_
from .._core import Module_
from ... import _ops


class PSDBlock_(Module_):
    """forward(x) = mfnet_(mean x) + fftnet_(x - mean x), log J = sum of the two (psd_.py:17-40)."""

    def __init__(self, *, mfnet_, fftnet_, label='psd-block'):
        super().__init__(label=label)
        self.mfnet_ = mfnet_
        self.fftnet_ = fftnet_

    # -- spectral evaluation (FFTNet_) ------------------------------------------------
    def _spectral(self, x, log0, inverse, parts=False):
        net = self.fftnet_
        V = int(np.prod(x.shape[1:]))
        rvol = float(V) ** 0.5
        ones = [1] * (x.dim() - 1)
        spec = net.spectrum(x)
        zero = torch.view_as_real(spec)[(slice(None),) + (0,) * (x.dim() - 1) + (0,)]
        x_mean = (zero / V).reshape(-1, *ones)
        mf = self.mfnet_.backward if inverse else self.mfnet_.forward
        y_mf, logJ_mf = mf(x_mean, rvol=rvol)
        w, logJ_fft = net.weights(inverse)
        logJ_fft = net.create_density(logJ_fft)
        if parts:
            y_fft = net.field(_ops.psd_scale(spec, w, torch.zeros_like(y_mf).reshape(-1), 0.0))
            return x_mean, (y_mf, logJ_mf), (y_fft, logJ_fft)
        y = net.field(_ops.psd_scale(spec, w, y_mf.reshape(-1), float(V)))
        return y, log0 + logJ_mf + logJ_fft

    # -- literal composition for any other pair of nets -------------------------------
    def _generic(self, x, log0, inverse):
        V = int(np.prod(x.shape[1:]))
        rvol = float(V) ** 0.5
        mean = _ops.sample_mean(x)
        x_mean = mean.reshape(-1, *[1] * (x.dim() - 1))
        mf = self.mfnet_.backward if inverse else self.mfnet_.forward
        ff = self.fftnet_.backward if inverse else self.fftnet_.forward
        y_mf, logJ_mf = mf(x_mean, rvol=rvol)
        y_fft, logJ_fft = ff(_ops.sample_shift(x, -mean))
        return _ops.sample_shift(y_fft, y_mf.reshape(-1)), log0 + logJ_mf + logJ_fft

    def _run(self, x, log0, inverse):
        if isinstance(self.fftnet_, FFTNet_) and x.dim() == 1 + self.fftnet_.lat_ndim:
            return self._spectral(x, log0, inverse)
        return self._generic(x, log0, inverse)

    def forward(self, x, log0=0):
        return self._run(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._run(x, log0, inverse=True)

    def _hack(self, x, log0=0):
        """Forward pass that also returns the intermediate parts (psd_.py:42-52)."""
        x_mean, (y_mf, logJ_mf), (y_fft, logJ_fft) = self._spectral(x, log0, False, parts=True)
        return [(x_mean, log0), (y_mf, logJ_mf), (y_fft, logJ_fft),
                (y_mf + y_fft, log0 + logJ_mf + logJ_fft)]

    def transfer(self, **kwargs):
        return self.__class__(mfnet_=self.mfnet_.transfer(**kwargs), fftnet_=self.fftnet_.transfer(**kwargs))
