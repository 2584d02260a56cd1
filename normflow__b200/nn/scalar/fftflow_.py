"""FFTNet_: rescale the Fourier modes of a field by 1/sqrt(ipsd(k^2)) (reference
src/nn/scalar/fftflow_.py).

The inverse power spectral density `ipsd` is a trainable monotone spline of the
normalised lattice momentum k^2 (IPSD, a SplineNet with two extra scale parameters).
The real-to-complex / complex-to-real transforms are cuFFT (through torch.fft); the
weights w = ipsd^(-1/2), their log-Jacobian and the multiplication of the half-spectrum
run in the package's own kernels (`nfk_psd_*`), which also give PSDBlock_ its fused
zero-mode handling.
"""

import copy

import numpy as np
import torch

from .modules import SplineNet
from .._core import Module_
from ... import _ops


def lattice_k2(lat_shape):
    """Lattice momentum squared sum_mu 4 sin^2(k_mu / 2), k_mu = 2 pi n / L_mu, on the rfftn
    half-spectrum (last axis trimmed to L/2 + 1) -- FreeScalar.calc_lattice_k2 /
    outer_lattice_k2, fftflow_.py:317-349.  float64 numpy."""
    total = np.zeros(())
    for n in lat_shape:
        k = np.linspace(0, 2 * np.pi * (1 - 1 / n), n)
        total = np.add.outer(total, 4 * np.sin(k / 2) ** 2)
    return total[..., :(1 + lat_shape[-1] // 2)]


class FreeScalar:
    """Free-theory helper (fftflow_.py:317-328)."""

    def __init__(self, lat_shape, kappa=None, m_sq=None):
        self.lat_shape = lat_shape
        self.kappa = kappa
        self.m_sq = m_sq

    def calc_lattice_k2(self):
        return torch.tensor(lattice_k2(self.lat_shape), dtype=torch.get_default_dtype())


class FFTNet_(Module_):
    """y = irfftn(rfftn(x) * w), w = 1/sqrt(ipsd), log J = sum over the full spectrum of
    log w (fftflow_.py:37-131,167-180).  The data may or may not carry a batch axis.

    lat_shape       : lattice shape (last axis even)
    ipsd_net        : module mapping the normalised k^2 grid to the inverse PSD
    ignore_zeromode : kept for `transfer`; the zero-mode convention lives in ipsd_net
    """

    def __init__(self, lat_shape, ipsd_net, ignore_zeromode=False, label='fftnet_'):
        super().__init__(label=label)
        lat_shape = tuple(int(n) for n in lat_shape)
        if lat_shape[-1] % 2 != 0:
            raise ValueError("FFTNet_: the last lattice axis must be even (the reference's irfftn "
                             "call returns a shorter axis otherwise)")
        self.lat_ndim = len(lat_shape)
        self.lat_shape = lat_shape
        self.ipsd_net = ipsd_net
        self.ignore_zeromode = ignore_zeromode
        self.rfft_dim = list(range(-self.lat_ndim, 0, 1))
        self.rfft_axis = -1
        k2 = lattice_k2(lat_shape)
        kmax = float(k2.max()) if k2.max() > 0 else 1.0
        self.register_buffer('norm_lat_k2', torch.tensor(k2 / kmax, dtype=torch.float32))
        self.register_buffer('max_lat_k2', torch.tensor(k2.max(), dtype=torch.float32))

    # ------------------------------------------------------------------ pieces used by PSDBlock_
    @property
    def ipsd(self):
        """Inverse power spectral density on the half-spectrum grid."""
        return self.ipsd_net(self.norm_lat_k2)

    def weights(self, inverse=False):
        """(w, log J) of the forward (or inverse) map."""
        return _ops.psd_weights(self.ipsd, inverse=inverse)

    def _check(self, x):
        if tuple(x.shape[-self.lat_ndim:]) != self.lat_shape:
            raise ValueError(f"FFTNet_ built for lattice {self.lat_shape}, got a field of shape {tuple(x.shape)}")

    def spectrum(self, x):
        self._check(x)
        return torch.fft.rfftn(x, dim=self.rfft_dim)

    def field(self, spec):
        return torch.fft.irfftn(spec, s=self.lat_shape, dim=self.rfft_dim)

    def _transform(self, x, log0, inverse):
        w, logj = self.weights(inverse)
        y = self.field(_ops.psd_scale(self.spectrum(x), w))
        return y, log0 + self.create_density(logj)

    # ------------------------------------------------------------------ Module_ protocol
    def forward(self, x, log0=0):
        return self._transform(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._transform(x, log0, inverse=True)

    def log_jacobian(self, weights):
        """log-Jacobian of multiplying the half-spectrum by `weights` (fftflow_.py:167-180): every
        mode counts twice (k and -k) except the self-conjugate planes at the two ends of the
        last axis."""
        dim = self.rfft_dim

        def sumlog(w):
            return torch.sum(torch.log(w), dim=dim)
        logj = 2 * sumlog(weights) - sumlog(weights[..., 0:1]) - sumlog(weights[..., -1:])
        return self.create_density(logj)

    def create_density(self, logJ):
        if Module_.propagate_density:
            n = int(np.prod(self.lat_shape))
            return (logJ.unsqueeze(-1) / n).expand(*logJ.shape, n).reshape(*logJ.shape, *self.lat_shape)
        return logJ

    @property
    def infrared_mass(self):
        """Dimensionless mass in lattice units."""
        return self.ipsd_net.infrared_mass(self.max_lat_k2)

    @staticmethod
    def build(lat_shape, knots_len=10, eff_mass2=1, eff_kappa=1, a=1, ignore_zeromode=False,
              nozeromode=False, **ipsd_kwargs):
        """FFTNet_ with an IPSD spline of `knots_len` knots started at the free theory
        eff_mass2 + eff_kappa k^2 (fftflow_.py:138-165)."""
        max_lat_k2 = float(lattice_k2(lat_shape).max())
        if knots_len < 2:     # two knots + smooth derivatives = identity spline
            knots_len = 2
            ipsd_kwargs.update(dict(smooth=True))
        logm2 = float(np.log(eff_mass2))
        logk2 = float(np.log(eff_kappa * max_lat_k2))
        scale = dict(a=a, ndim=len(lat_shape))
        if nozeromode and not ignore_zeromode:
            logy = IPSDnozeromode.apply_scale(torch.tensor([logk2]), **scale)
            ipsd_net = IPSDnozeromode(knots_len, logy=logy, **ipsd_kwargs)
        else:
            logy = IPSD.apply_scale(torch.tensor([logm2, logk2]), **scale)
            ipsd_net = IPSD(knots_len, logy=logy, ignore_zeromode=ignore_zeromode, **ipsd_kwargs)
        return FFTNet_(lat_shape, ipsd_net, ignore_zeromode=ignore_zeromode)

    def transfer(self, scale_factor=1, shape=None, **extra):
        """The same spectral density on another lattice: `shape` is the new lattice, `scale_factor`
        the ratio old/new lattice spacing (fftflow_.py:187-212)."""
        shape = self.lat_shape if shape is None else shape
        ipsd_net = self.ipsd_net.transfer(scale_factor=scale_factor, ndim=self.lat_ndim)
        return self.__class__(shape, ipsd_net=ipsd_net, ignore_zeromode=self.ignore_zeromode)


def _pin_first(t, value):
    """t[0, ..., 0] = value in place.  A fill on a view: unlike indexed assignment of a python
    number it stages no host scalar, so it can be captured into a CUDA graph."""
    t.view(-1)[:1].fill_(value)


class _ScaledSpline(SplineNet):
    """Shared by the two IPSD flavours: deep copy with rescaled `logy`."""

    def transfer(self, scale_factor=1, ndim=1):
        ipsd = copy.deepcopy(self)
        with torch.no_grad():
            new = self.apply_scale(self.logy.detach(), a=1 / scale_factor, ndim=ndim)
            ipsd.logy.copy_(new.to(ipsd.logy.device))
        return ipsd


class IPSD(_ScaledSpline):
    """Inverse power spectral density  e^{logy0} + e^{logy1} * spline(k^2 / k^2_max)
    (fftflow_.py:226-267).  With ignore_zeromode the k = 0 entry is pinned to 1, which
    removes the zero mode from the Jacobian."""

    def __init__(self, knots_len, *, logy, ignore_zeromode=False, **kwargs):
        super().__init__(knots_len, **kwargs)
        self.logy = torch.nn.Parameter(torch.as_tensor(logy, dtype=torch.float32).clone())
        self.ignore_zeromode = ignore_zeromode

    def forward(self, x):
        y = torch.exp(self.logy)
        sigma_k2 = y[0] + y[1] * super().forward(x)
        if self.ignore_zeromode:
            _pin_first(sigma_k2, 1.0)
        return sigma_k2

    def _backward(self, x):
        y = torch.exp(self.logy)
        return super().backward((x - y[0]) / y[1])

    @staticmethod
    @torch.no_grad()
    def apply_scale(logy, *, a, ndim):
        loga = float(np.log(a))
        return torch.stack([logy[0] + loga * ndim, logy[1] + loga * (ndim - 2)])

    @torch.no_grad()
    def infrared_mass(self, max_lat_k2):
        return torch.exp(0.5 * self.logy[0])


class IPSDnozeromode(_ScaledSpline):
    """e^{logy0} * spline(k^2 / k^2_max) with the k = 0 entry pinned to 1
    (fftflow_.py:270-313; the reference prefers IPSD(ignore_zeromode=True))."""

    def __init__(self, knots_len, *, logy, **kwargs):
        super().__init__(knots_len, **kwargs)
        self.logy = torch.nn.Parameter(torch.as_tensor(logy, dtype=torch.float32).clone())

    def forward(self, x):
        sigma_k2 = torch.exp(self.logy)[0] * super().forward(x)
        _pin_first(sigma_k2, 1.0)
        return sigma_k2

    def _backward(self, x):
        x = x / torch.exp(self.logy)[0]
        _pin_first(x, 0.0)
        return super().backward(x)

    @staticmethod
    @torch.no_grad()
    def apply_scale(logy, *, a, ndim):
        return (logy[0] + float(np.log(a)) * (ndim - 2)).reshape(1)

    @torch.no_grad()
    def infrared_mass(self, max_lat_k2):
        probe = torch.tensor([1e-6, 2e-6], dtype=torch.float32, device=self.logy.device) / max_lat_k2
        z = self.forward(probe)
        factor = (z[1] - z[0]) / 1e-6
        return (z[0] / factor) ** 0.5
