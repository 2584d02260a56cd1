"""Lattice actions (reference src/action/scalar_action.py)."""

import torch

from .. import _ops


class ScalarPhi4Action:
    r"""phi^4 theory on a periodic lattice,

        S = sum_x [ kappa/2 (d_mu phi)^2 + m^2/2 phi^2 + lambda phi^4 ].

    `action` runs one fused stencil kernel (nearest-neighbour products, powers and the
    per-sample reduction in a single pass; gradient kernel for autograd) instead of the
    reference's D rolled copies and D+1 reductions (scalar_action.py:40-46).
    """

    def __init__(self, *, m_sq, lambd, kappa=1, a=1):
        self.kappa, self.m_sq, self.lambd, self.a = kappa, m_sq, lambd, a

    def get_coef(self, lat_ndim):
        """(w0, w2, w4) with the lattice spacing absorbed (scalar_action.py:22-33):
        S = sum w2 phi^2 + w4 phi^4 - w0 sum_mu phi(x) phi(x - mu)."""
        a = self.a
        kappa = self.kappa * a ** (lat_ndim - 2)
        w_0 = kappa
        w_2 = 0.5 * (self.m_sq * a ** lat_ndim + 2 * kappa * lat_ndim)
        w_4 = self.lambd * a ** lat_ndim
        return w_0, w_2, w_4

    def __call__(self, cfgs):
        return self.action(cfgs)

    def action(self, cfgs):
        """S[b] for a batch of configurations (B, *lattice)."""
        w0, w2, w4 = self.get_coef(cfgs.ndim - 1)
        return _ops.phi4_action(cfgs, w0, w2, w4)

    def action_density(self, cfgs):
        """Symmetric, positive-kinetic-term density (scalar_action.py:48-62).  A
        diagnostic, not on the hot path: plain tensor expressions."""
        dims = tuple(range(1, cfgs.ndim))
        w0, w2, w4 = self.get_coef(cfgs.ndim - 1)
        dens = (w2 - w0 * len(dims)) * cfgs ** 2 + w4 * cfgs ** 4
        for mu in dims:
            for step in (-1, +1):
                dens = dens + (w0 / 4) * (cfgs - torch.roll(cfgs, step, mu)) ** 2
        return dens

    def potential(self, x):
        return self.m_sq * x ** 2 + self.lambd * x ** 4

    def log_prob(self, x, action_logz=0):
        """log probability up to an additive constant."""
        return -self.action(x) - action_logz
