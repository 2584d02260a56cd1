"""Site-local invertible layers (reference src/nn/scalar/modules_.py).

DistConvertor_ -- the chain Expit_ -> SplineNet_ -> Logit_ -- runs as ONE kernel that
carries every point of (0,1) as the pair (s, 1-s), so the tails keep float32 relative
accuracy, and reduces the three log-Jacobians per sample in the same pass.
"""

import copy

import numpy as np
import torch

from .modules import SplineNet
from .._core import Module_, ModuleList_
from ... import _ops


class Identity_(Module_):

    def __init__(self, label='identity_'):
        super().__init__(label=label)

    def forward(self, x, log0=0, **extra):
        return x, log0

    def backward(self, x, log0=0, **extra):
        return x, log0


class Clone_(Module_):

    def __init__(self, label='clone_'):
        super().__init__(label=label)

    def forward(self, x, log0=0, **extra):
        return x.clone(), log0

    def backward(self, x, log0=0, **extra):
        return x.clone(), log0


class ScaleNet_(Module_):
    """x -> w x with one positive trainable w = softplus_{ln 2}(_weight)
    (reference modules_.py:44-69)."""

    def __init__(self, label='scale_'):
        super().__init__(label=label)
        self._weight = torch.nn.Parameter(torch.zeros(1))

    @property
    def weight(self):
        return torch.nn.functional.softplus(self._weight, beta=float(np.log(2)))

    def _logj(self, x):
        nvar = int(np.prod(x.shape[1:]))
        return torch.log(self.weight) * nvar * torch.ones(x.shape[0], device=self._weight.device)

    def forward(self, x, log0=0):
        return x * self.weight, log0 + self._logj(x)

    def backward(self, x, log0=0):
        return x / self.weight, log0 - self._logj(x)


class Expit_(Module_):
    """y = 1/(1+e^-x), log J = sum(-x + 2 log y)  (reference modules_.py:93-102)."""

    def forward(self, x, log0=0):
        return _ops.logistic(x, 0, log0)

    def backward(self, x, log0=0):
        return _ops.logistic(x, 1, log0)


class Logit_(Module_):
    """y = log(x/(1-x)), log J = -sum log(x (1-x))  (reference modules_.py:105-114)."""

    def forward(self, x, log0=0):
        return _ops.logistic(x, 1, log0)

    def backward(self, x, log0=0):
        return _ops.logistic(x, 0, log0)


class SplineNet_(SplineNet, Module_):
    """SplineNet that also reports log|dy/dx| summed per sample
    (reference modules_.py:277-302)."""

    def _extrap(self):
        return self.spline_kwargs.get('extrap', {})

    def forward(self, x, log0=0):
        return _ops.spline1d(x, self.knots(), log0, self._extrap())

    def backward(self, x, log0=0):
        return _ops.spline1d(x, self.knots(), log0, self._extrap(), inverse=True)


class UnityDistConvertor_(SplineNet_):
    """Distribution convertor for variables in [0, 1] (reference modules_.py:305-316)."""

    def __init__(self, knots_len, symmetric=False, **kwargs):
        extra = dict(xlim=(0.5, 1), ylim=(0.5, 1), extrap={'left': 'anti'}) if symmetric else {}
        super().__init__(knots_len, **kwargs, **extra)


class PhaseDistConvertor_(SplineNet_):
    """Distribution convertor for variables in [-pi, pi] (reference modules_.py:319-330)."""

    def __init__(self, knots_len, symmetric=False, label='phase-dc_', **kwargs):
        pi = np.pi
        if symmetric:
            extra = dict(xlim=(0, pi), ylim=(0, pi), extrap={'left': 'anti'})
        else:
            extra = dict(xlim=(-pi, pi), ylim=(-pi, pi))
        super().__init__(knots_len, label=label, **kwargs, **extra)


class SgnBiasNet_(Module_):
    """x -> x + sgn(x) w^2; only meaningful as the very first layer
    (reference modules_.py:386-400)."""

    def __init__(self, size=[1], label='sgnbias_'):
        super().__init__(label=label)
        self.w = torch.nn.Parameter(torch.rand(*size) / 10)

    def forward(self, x, log0=0):
        return x + torch.sgn(x) * self.w ** 2, log0

    def backward(self, x, log0=0):
        return x - torch.sgn(x) * self.w ** 2, log0


class DistConvertor_(ModuleList_):
    """Distribution convertor for real variables: Expit_, SplineNet_, Logit_
    (reference modules_.py:333-383).  symmetric=True makes the map odd (spline on
    [0.5, 1] with an anti-periodic left boundary).

    The three layers stay in the list (labels, parameters and state_dict keys as in the
    reference: `1.weights_x` ...), but forward/backward evaluate the chain in a single
    fused kernel whenever it is exactly [Expit_, SplineNet_, Logit_]; optional
    ScaleNet_ / SgnBiasNet_ layers run around it.
    """

    def __init__(self, knots_len, symmetric=False, label='dc_', sgnbias=False, initial_scale=False,
                 final_scale=False, **kwargs):
        if symmetric:
            extra = dict(xlim=(0.5, 1), ylim=(0.5, 1), extrap={'left': 'anti'})
        else:
            extra = dict(xlim=(0, 1), ylim=(0, 1))
        if knots_len > 1:
            spline_ = SplineNet_(knots_len, label='spline_', **kwargs, **extra)
            nets_ = [Expit_(label='expit_'), spline_, Logit_(label='logit_')]
        else:
            nets_ = []
        if initial_scale:
            nets_ = [ScaleNet_(label='scale_')] + nets_
        elif final_scale:
            nets_ = nets_ + [ScaleNet_(label='scale_')]
        if sgnbias:      # must come first if present
            nets_ = [SgnBiasNet_()] + nets_
        super().__init__(nets_)
        self.label = label
        self.symmetric = symmetric

    def transfer(self, **kwargs):
        """A pointwise map does not depend on the lattice: a deep copy.  (The reference inherits
        ModuleList_.transfer here, which calls DistConvertor_(list_of_layers) and raises TypeError.)"""
        return copy.deepcopy(self)

    def _layer(self, label):
        for net_ in self:
            if net_.label == label:
                return net_
        return None

    @property
    def spline_layer_(self):
        return self._layer('spline_')

    @property
    def scale_layer_(self):
        return self._layer('scale_')

    @property
    def sgnbias_layer_(self):
        return self._layer('sgnbias_')

    def _segments(self):
        """Split the list into runs; a run [Expit_, SplineNet_, Logit_] becomes 'chain'."""
        nets = list(self)
        out, i = [], 0
        while i < len(nets):
            if (i + 2 < len(nets) and isinstance(nets[i], Expit_) and isinstance(nets[i + 1], SplineNet_)
                    and isinstance(nets[i + 2], Logit_)):
                out.append(('chain', nets[i + 1]))
                i += 3
            else:
                out.append(('layer', nets[i]))
                i += 1
        return out

    def _chain(self, spline_, x, log0, inverse):
        return _ops.spline1d(x, spline_.knots(both_ends=True), log0, spline_._extrap(),
                             logistic_wrap=True, inverse=inverse)

    def forward(self, x, log0=0):
        for kind, net_ in self._segments():
            if kind == 'chain':
                x, log0 = self._chain(net_, x, log0, inverse=False)
            else:
                x, log0 = net_.forward(x, log0)
        return x, log0

    def backward(self, x, log0=0):
        for kind, net_ in reversed(self._segments()):
            if kind == 'chain':
                x, log0 = self._chain(net_, x, log0, inverse=True)
            else:
                x, log0 = net_.backward(x, log0)
        return x, log0
