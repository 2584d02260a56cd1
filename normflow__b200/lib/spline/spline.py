"""Rational-quadratic (Pade [2,2]) spline with shared 1-D knots
(reference src/lib/spline/spline.py: SplineTemplate :39-123, Pade22Spline :175-287,
AugmentKnots :392-540).

The evaluation -- bin search, segment value and derivative, extrapolation by a
straight line, the anti-periodic or the periodic mirror image -- is one kernel
(`nfk_spline1d_*`); extra fiducial knots are never materialised.  Per-site knots
(a knot tensor with batch / lattice axes) exist only inside the coupling kernels,
see nn.RQSplineCoupling_.
"""

import torch

from ... import _ops


class RQSpline:
    """Monotone C^1 interpolant through (knots_x, knots_y) with derivatives knots_d.

    knots_* : float32[K] tensors (K >= 2).  knots_d=None -> each knot takes the mean
              slope of its two segments, end knots the slope of their own segment.
    extrap  : dict(left=..., right=...) with None (end segment extended), 'linear',
              'anti' / 'anti-periodic' (point mirror about the end knot) or 'periodic' (even mirror
              image; the end knot's derivative must be zero; forward direction only -- the map is
              not monotone there and the derivative returned with grad=True is negative).
    """

    def __init__(self, knots_x=None, knots_y=None, knots_d=None, knots_axis=-1, extrap={}):
        if knots_x.dim() != 1 or knots_y.dim() != 1:
            raise NotImplementedError("RQSpline: shared 1-D knots only; per-site knots live in "
                                      "RQSplineCoupling_'s fused kernel")
        if knots_d is None:
            knots_d = self.smooth_derivatives(knots_x, knots_y)
        for side, end in (('left', 0), ('right', -1)):
            if extrap.get(side) == 'periodic' and float(knots_d[end]) != 0.0:
                raise Exception("Oops: derivative at periodic bc must be zero.")    # spline.py:504-505, 520-521
        self.knots_x, self.knots_y, self.knots_d = knots_x, knots_y, knots_d
        self.knots_axis = knots_axis
        self.extrap = dict(extrap)
        self.knots_len = knots_x.shape[0]
        self.segm_len = self.knots_len - 1

    @staticmethod
    def smooth_derivatives(knots_x, knots_y, knots_axis=-1):
        slope = (knots_y[1:] - knots_y[:-1]) / (knots_x[1:] - knots_x[:-1])
        return torch.cat([slope[:1], 0.5 * (slope[1:] + slope[:-1]), slope[-1:]])

    def __call__(self, x, **kwargs):
        return self.forward(x, **kwargs)

    def _table(self):
        packed = getattr(self, '_packed', None)
        return packed if packed is not None else torch.stack([self.knots_x, self.knots_y, self.knots_d])

    def _run(self, x, grad, inverse):
        flat = x.reshape(1, -1)
        y, _ = _ops.spline1d(flat, self._table(), 0, self.extrap, logistic_wrap=False, inverse=inverse)
        y = y.reshape(x.shape)
        if not grad:
            return y
        return y, self._derivative(x if not inverse else y, inverse)

    def _derivative(self, x, inverse):
        """dy/dx at x (or its reciprocal for the inverse map) as a tensor: evaluated by
        autograd through the forward kernel's own backward kernel."""
        with torch.enable_grad():
            xr = x.detach().reshape(1, -1).requires_grad_(True)
            y, _ = _ops.spline1d(xr, self._table().detach(), 0, self.extrap, logistic_wrap=False)
            (g,) = torch.autograd.grad(y.sum(), xr)
        g = g.reshape(x.shape)
        return 1.0 / g if inverse else g

    def forward(self, x, grad=False, squeezed=False):
        """Spline value at x; with grad=True also dy/dx."""
        return self._run(x, grad, inverse=False)

    def backward(self, y, grad=False, squeezed=False):
        """Inverse map; with grad=True also dx/dy."""
        if 'periodic' in (self.extrap.get('left'), self.extrap.get('right')):
            raise NotImplementedError("RQSpline.backward: a periodically continued spline is not monotone "
                                      "(the reference's searchsorted over its mirrored knots_y is undefined too)")
        return self._run(y, grad, inverse=True)


Pade22Spline = RQSpline
