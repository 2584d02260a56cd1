"""CPU oracle for the normflow hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy float64 restatement of the reference's algorithm for the
`Model.fit / posterior.sample / mcmc.sample` inner loop (NormalPrior log-prob,
checkerboard masks, affine / shift / rational-quadratic-spline couplings with
log|det J|, DistConvertor_, circular ConvAct conditioner, phi^4 action,
Metropolis accept/reject) and of the PSDBlock_ that opens examples/scalar_affine.py
(MeanFieldNet_ + FFTNet_/IPSD).  Every function cites the reference file:line it
follows (paths relative to /root/reference/).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this module, and only as the checker or
as the reported CPU baseline.  Nothing under `normflow__b200/` imports it.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, imported from
/root/reference in the build container by `tests/golden/make_golden.py`; the
resulting fixtures are committed under `tests/golden/*.npz` and checked by
`tests/test_oracle_golden.py` (CPU, no GPU needed), together with the known-answer
values of SURVEY.md section 4.

All functions are dtype-generic in the sense that complex128 inputs are accepted
where the maths is analytic: `directional_derivative` uses that (complex-step
differentiation) to give derivative checks that are independent of the
hand-derived adjoints in the CUDA kernels.
"""

import itertools

import numpy as np

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)
LN2 = np.log(2.0)


def _re(z):
    """Real part used for every ordering / branching decision."""
    return z.real if np.iscomplexobj(z) else z


# =============================================================================
# masks  (src/mask/mask.py)
# =============================================================================
def evenodd_mask(shape, parity=0, exclude_mu=None):
    """EvenOddMask.make_mask, src/mask/mask.py:53-61 (bit-exact contract).

    mask[ind] = (1 - parity + sum(ind) [- ind[exclude_mu]]) % 2, uint8.
    Written as the same per-site loop as the reference (python `%` semantics).
    """
    shape = tuple(int(l) for l in shape)
    mask = np.empty(shape, dtype=np.uint8)
    for ind in itertools.product(*[range(l) for l in shape]):
        s = sum(ind) if exclude_mu is None else sum(ind) - ind[exclude_mu]
        mask[ind] = (1 - parity + s) % 2
    return mask


def alongaxes_mask(shape, parity=0, mu=0):
    """AlongAxesEvenOddMask.make_mask, src/mask/mask.py:64-72."""
    shape = tuple(int(l) for l in shape)
    mask = np.empty(shape, dtype=np.uint8)
    for ind in itertools.product(*[range(l) for l in shape]):
        mask[ind] = (1 - parity + ind[mu]) % 2
    return mask


def mask_split(mask, x):
    """Mask.split, src/mask/mask.py:30-31: (m*x, (1-m)*x)."""
    return mask * x, (1 - mask) * x


def mask_purify(mask, x, channel):
    """Mask.purify, src/mask/mask.py:36-37."""
    return x * (mask if channel == 0 else (1 - mask))


# =============================================================================
# prior  (src/prior/prior.py)
# =============================================================================
def normal_log_prob(x, loc=0.0, scale=1.0):
    """Prior.log_prob for NormalPrior, src/prior/prior.py:30-36,92-104.

    torch.distributions.Normal.log_prob = -(x-loc)^2/(2 s^2) - log s - log sqrt(2 pi),
    summed over every axis but the batch axis.
    """
    dens = -((x - loc) ** 2) / (2.0 * scale ** 2) - np.log(scale) - LOG_SQRT_2PI
    dens = dens + np.zeros_like(x)  # broadcast scalars
    return dens.reshape(x.shape[0], -1).sum(axis=1)


# =============================================================================
# action  (src/action/scalar_action.py)
# =============================================================================
def phi4_coef(ndim, *, m_sq, lambd, kappa=1.0, a=1.0):
    """ScalarPhi4Action.get_coef, src/action/scalar_action.py:22-33."""
    kap = kappa * a ** (ndim - 2)
    w0 = 0.5 * (2 * kap)
    w2 = 0.5 * (m_sq * a ** ndim + 2 * kap * ndim)
    w4 = lambd * a ** ndim
    return w0, w2, w4


def phi4_action(cfgs, *, m_sq, lambd, kappa=1.0, a=1.0):
    """ScalarPhi4Action.action, src/action/scalar_action.py:38-46.

    S[b] = sum_x (w2 phi^2 + w4 phi^4) - w0 sum_mu sum_x phi(x) phi(x - mu)
    with periodic wrap (torch.roll(cfgs, 1, mu)).
    """
    axes = tuple(range(1, cfgs.ndim))
    w0, w2, w4 = phi4_coef(cfgs.ndim - 1, m_sq=m_sq, lambd=lambd, kappa=kappa, a=a)
    act = np.sum(w2 * cfgs ** 2 + w4 * cfgs ** 4, axis=axes)
    for mu in axes:
        act = act - w0 * np.sum(cfgs * np.roll(cfgs, 1, axis=mu), axis=axes)
    return act


def phi4_action_grad(cfgs, *, m_sq, lambd, kappa=1.0, a=1.0):
    """dS/dphi(x) = 2 w2 phi + 4 w4 phi^3 - w0 sum_mu (phi(x+mu) + phi(x-mu));
    the derivative of scalar_action.py:38-46 (what autograd produces there)."""
    axes = tuple(range(1, cfgs.ndim))
    w0, w2, w4 = phi4_coef(cfgs.ndim - 1, m_sq=m_sq, lambd=lambd, kappa=kappa, a=a)
    g = 2 * w2 * cfgs + 4 * w4 * cfgs ** 3
    for mu in axes:
        g = g - w0 * (np.roll(cfgs, 1, axis=mu) + np.roll(cfgs, -1, axis=mu))
    return g


def phi4_action_density(cfgs, *, m_sq, lambd, kappa=1.0, a=1.0):
    """ScalarPhi4Action.action_density, src/action/scalar_action.py:48-62."""
    axes = tuple(range(1, cfgs.ndim))
    nd = cfgs.ndim - 1
    w0, w2, w4 = phi4_coef(nd, m_sq=m_sq, lambd=lambd, kappa=kappa, a=a)
    w2 = w2 - w0 * nd
    dens = w2 * cfgs ** 2 + w4 * cfgs ** 4
    for mu in axes:
        dens = dens + (w0 / 4) * (cfgs - np.roll(cfgs, -1, axis=mu)) ** 2
        dens = dens + (w0 / 4) * (cfgs - np.roll(cfgs, +1, axis=mu)) ** 2
    return dens


# =============================================================================
# spline parameterisation  (src/nn/scalar/couplings_.py:211-262, modules.py:369-391)
# =============================================================================
def softplus_ln2(z, threshold=20.0):
    """torch.nn.Softplus(beta=ln 2): log2(1 + 2^z), linear when beta*z > 20
    (couplings_.py:173-176)."""
    bz = LN2 * z
    safe = np.where(_re(bz) > threshold, 0.0, bz)
    return np.where(_re(bz) > threshold, z, np.log1p(np.exp(safe)) / LN2)


def softmax(w, axis):
    """torch.nn.Softmax(dim=axis) (max-subtracted)."""
    wmax = np.max(_re(w), axis=axis, keepdims=True)
    e = np.exp(w - wmax)
    return e / np.sum(e, axis=axis, keepdims=True)


def to_coord(w, axis):
    """cat(0, cumsum(softmax(w))) along `axis`  (couplings_.py:235-236)."""
    c = np.cumsum(softmax(w, axis), axis=axis)
    pad_shape = list(c.shape)
    pad_shape[axis] = 1
    return np.concatenate([np.zeros(pad_shape, dtype=c.dtype), c], axis=axis)


def knots_from_raw(out, *, xlim, ylim, axis=1, knots_x=None, knots_y=None):
    """RQSplineCoupling_.make_spline, src/nn/scalar/couplings_.py:211-262: split the
    conditioner output into (K-1, K-1, K) channels; with fixed 1-D `knots_x` and / or
    `knots_y` (:246-258) into (K-1, K) or K channels -- the fixed coordinate is used
    as it stands, broadcast over batch and lattice."""
    n = out.shape[axis]
    if knots_x is not None or knots_y is not None:
        def fixed(k):
            shape = [1] * out.ndim
            shape[axis] = -1
            full = list(out.shape)
            full[axis] = len(k)
            return np.broadcast_to(np.asarray(k, dtype=np.float64).reshape(shape), full)
        if knots_x is not None and knots_y is not None:
            return fixed(knots_x), fixed(knots_y), softplus_ln2(out)
        m = (n + 2) // 2
        assert 2 * m - 1 == n, "conditioner must emit 2K-1 channels when one coordinate is fixed"
        free, d_ = np.split(out, [m - 1], axis=axis)
        if knots_x is not None:
            return fixed(knots_x), to_coord(free, axis) * (ylim[1] - ylim[0]) + ylim[0], softplus_ln2(d_)
        return to_coord(free, axis) * (xlim[1] - xlim[0]) + xlim[0], fixed(knots_y), softplus_ln2(d_)
    m = (n + 2) // 3
    assert 3 * m - 2 == n, "conditioner must emit 3K-2 channels"
    x_, y_, d_ = np.split(out, [m - 1, 2 * (m - 1)], axis=axis)
    kx = to_coord(x_, axis) * (xlim[1] - xlim[0]) + xlim[0]
    ky = to_coord(y_, axis) * (ylim[1] - ylim[0]) + ylim[0]
    kd = softplus_ln2(d_)
    return kx, ky, kd


def smooth_derivatives(kx, ky, axis):
    """SplineTemplate.smooth_derivatives (bc_type='not-ones'),
    src/lib/spline/spline.py:125-152."""
    kx = np.moveaxis(kx, axis, -1)
    ky = np.moveaxis(ky, axis, -1)
    m = (ky[..., 1:] - ky[..., :-1]) / (kx[..., 1:] - kx[..., :-1])
    m_avg = 0.5 * (m[..., 1:] + m[..., :-1])
    d = np.concatenate([m[..., :1], m_avg, m[..., -1:]], axis=-1)
    return np.moveaxis(d, -1, axis)


# =============================================================================
# rational-quadratic spline  (src/lib/spline/spline.py)
# =============================================================================
class RQSpline:
    """Pade22Spline / RQSpline, src/lib/spline/spline.py:39-68, 87-123, 154-287,
    including AugmentKnots (:392-540).  knots_* are arrays with the knots along
    `axis` (same shape), or 1-D (shared knots).
    """

    def __init__(self, knots_x, knots_y, knots_d=None, axis=-1, extrap=None,
                 stable_inverse=False):
        extrap = dict(extrap or {})
        one_dim = (knots_x.ndim == 1)
        if one_dim:
            axis = -1
        if knots_d is None:
            knots_d = smooth_derivatives(knots_x, knots_y, axis)
        kx, ky, kd = self._augment(knots_x, knots_y, knots_d, axis,
                                   extrap.get('left'), extrap.get('right'))
        self.kx, self.ky, self.kd = kx, ky, kd
        self.axis = axis
        self.one_dim = one_dim
        self.n_seg = kx.shape[axis] - 1
        self.stable_inverse = stable_inverse

    # --- AugmentKnots, spline.py:392-540 -------------------------------------
    @staticmethod
    def _augment(x, y, d, axis, left, right):
        take = lambda z, sl: np.take(z, sl, axis=axis)
        flip = lambda z: np.flip(z, axis=axis)
        n = x.shape[axis]
        if left is None and right is None:
            return x, y, d
        if left == 'linear' or right == 'linear':
            # takecare_linear, spline.py:458-486: one fiducial knot one unit out
            parts_x, parts_y, parts_d = [x], [y], [d]
            if left == 'linear':
                parts_x.insert(0, take(x, [0]) - 1)
                parts_y.insert(0, take(y, [0]) - take(d, [0]))
                parts_d.insert(0, take(d, [0]))
            if right == 'linear':
                parts_x.append(take(x, [n - 1]) + 1)
                parts_y.append(take(y, [n - 1]) + take(d, [n - 1]))
                parts_d.append(take(d, [n - 1]))
            x = np.concatenate(parts_x, axis=axis)
            y = np.concatenate(parts_y, axis=axis)
            d = np.concatenate(parts_d, axis=axis)
            if left is None or right is None:
                return x, y, d  # perform_bc early return, spline.py:451-455
        # takecare_rest, spline.py:488-532 (acts on the possibly augmented knots)
        n = x.shape[axis]
        lo, hi = list(range(1, n)), list(range(0, n - 1))
        px, py, pd = [x], [y], [d]
        if left in ('anti', 'anti-periodic'):
            px.insert(0, 2 * take(x, [0]) - flip(take(x, lo)))
            py.insert(0, 2 * take(y, [0]) - flip(take(y, lo)))
            pd.insert(0, flip(take(d, lo)))
        elif left == 'periodic':
            px.insert(0, 2 * take(x, [0]) - flip(take(x, lo)))
            py.insert(0, flip(take(y, lo)))
            pd.insert(0, -flip(take(d, lo)))
        if right in ('anti', 'anti-periodic'):
            px.append(2 * take(x, [n - 1]) - flip(take(x, hi)))
            py.append(2 * take(y, [n - 1]) - flip(take(y, hi)))
            pd.append(flip(take(d, hi)))
        elif right == 'periodic':
            px.append(2 * take(x, [n - 1]) - flip(take(x, hi)))
            py.append(flip(take(y, hi)))
            pd.append(-flip(take(d, hi)))
        cat = lambda p: np.concatenate(p, axis=axis) if len(p) > 1 else p[0]
        return cat(px), cat(py), cat(pd)

    # --- searchsorted + clamp, spline.py:154-172 -----------------------------
    def _segment(self, sorted_knots, v):
        """torch.searchsorted(right=False): number of knots strictly below v;
        then clamp(., 1, n_seg) - 1."""
        if self.one_dim:
            idx = np.searchsorted(_re(sorted_knots), _re(v).ravel(), side='left')
            idx = idx.reshape(v.shape)
        else:
            idx = np.sum(_re(sorted_knots) < _re(v), axis=self.axis, keepdims=True)
        return np.clip(idx, 1, self.n_seg) - 1

    def _gather(self, knots, seg, shift):
        if self.one_dim:
            return knots[seg + shift]
        return np.take_along_axis(knots, seg + shift, axis=self.axis)

    def _segment_params(self, seg):
        x0, x1 = self._gather(self.kx, seg, 0), self._gather(self.kx, seg, 1)
        y0, y1 = self._gather(self.ky, seg, 0), self._gather(self.ky, seg, 1)
        d0, d1 = self._gather(self.kd, seg, 0), self._gather(self.kd, seg, 1)
        return x0, x1, y0, y1, d0, d1

    @staticmethod
    def _g1(theta, m, d0, d1):
        # spline.py:210-212
        return m ** 2 * (d0 + 2 * (m - d0) * theta + (d1 + d0 - 2 * m) * theta ** 2) \
            / (m + (d1 + d0 - 2 * m) * theta * (1 - theta)) ** 2

    def forward(self, x):
        """SplineTemplate.forward + Pade22Spline._calc_segment_func,
        spline.py:87-115, 185-220.  `x` has a size-1 axis at `axis` unless the
        knots are 1-D.  Returns (g0, g1)."""
        seg = self._segment(self.kx, x)
        x0, x1, y0, y1, d0, d1 = self._segment_params(seg)
        m = (y1 - y0) / (x1 - x0)
        theta = (x - x0) / (x1 - x0)
        g0 = y0 + (y1 - y0) * theta * (m * theta + d0 * (1 - theta)) \
            / (m + (d1 + d0 - 2 * m) * theta * (1 - theta))
        return g0, self._g1(theta, m, d0, d1)

    def backward(self, y):
        """SplineTemplate.backward + _calc_segment_inv_func, spline.py:117-123,
        222-287.  Returns (x, 1/g1).  With stable_inverse=False this follows the
        reference's root formula literally (ill-conditioned on linear fiducial
        segments, SURVEY.md section 7 hard part 3)."""
        seg = self._segment(self.ky, y)
        x0, x1, y0, y1, d0, d1 = self._segment_params(seg)
        m = (y1 - y0) / (x1 - x0)
        eta = (y - y0) / (y1 - y0)
        a2 = (2 * m - d1 - d0) * eta + d0 - m
        a1 = -a2 - m
        a0 = m * eta
        delta = np.sqrt(a1 ** 2 - 4 * a0 * a2)
        if self.stable_inverse:
            theta = 2 * a0 / (-a1 + delta)
        else:
            with np.errstate(divide='ignore', invalid='ignore'):
                quad = (-a1 - delta) / (2 * a2)
                lin = -a0 / a1
            theta = np.where(a2 == 0, lin, quad)
        x = x0 + (x1 - x0) * theta
        return x, 1.0 / self._g1(theta, m, d0, d1)


# =============================================================================
# conditioner: circular "same" convolution stack  (modules.py:68-154, convNd.py)
# =============================================================================
def conv_circular(inp, weight, bias=None):
    """torch.nn.Conv{1,2,3}d(padding='same', padding_mode='circular') and
    ConvNd/Conv4d (convNd.py:84-127): cross-correlation
        out[b,o,s] = bias[o] + sum_{i,k} w[o,i,k] * in[b,i,(s + k - k_size//2) mod L].
    inp (B,Ci,*L), weight (Co,Ci,*k) with odd k.
    """
    nd = inp.ndim - 2
    ksize = weight.shape[2:]
    assert len(ksize) == nd and all(k % 2 == 1 for k in ksize)
    out = np.zeros((inp.shape[0], weight.shape[0]) + inp.shape[2:],
                   dtype=np.result_type(inp.dtype, weight.dtype))
    for tap in itertools.product(*[range(k) for k in ksize]):
        shifted = inp
        for ax, (t, k) in enumerate(zip(tap, ksize)):
            shifted = np.roll(shifted, -(t - k // 2), axis=2 + ax)
        w_tap = weight[(slice(None), slice(None)) + tap]       # (Co, Ci)
        out += np.tensordot(w_tap, shifted, axes=([1], [1])).swapaxes(0, 1)
    if bias is not None:
        out += bias.reshape((1, -1) + (1,) * nd)
    return out


def conv4d_lower_to_standard(w_lower, out_channels, k0):
    """ConvNd._to_standard_weight_shape, convNd.py:129-142: the stored
    `_conv_lower_dim.weight` (Co*k0, Ci, k,k,k) -> standard (Co, Ci, k0,k,k,k)."""
    ci = w_lower.shape[1]
    rest = w_lower.shape[2:]
    w = w_lower.reshape((out_channels, k0, ci) + rest)
    return np.moveaxis(w, 1, 2)


_ACTS = {
    None: lambda v: v,
    'none': lambda v: v,
    'tanh': np.tanh,
    'relu': lambda v: np.where(_re(v) > 0, v, 0.0),
    'leaky_relu': lambda v: np.where(_re(v) > 0, v, 0.01 * v),
    'softplus': lambda v: np.where(_re(v) > 20, v, np.log1p(np.exp(np.where(_re(v) > 20, 0, v)))),
    'abs': lambda v: np.where(_re(v) >= 0, v, -v),
}


def convact_forward(inp, layers, acts, pre_act=None):
    """ConvAct, src/nn/scalar/modules.py:120-154: [pre_act] -> (conv -> act)*.
    `layers` = list of (weight, bias_or_None); `acts` = list of names/None."""
    h = _ACTS[pre_act](inp)
    for (w, b), act in zip(layers, acts):
        h = _ACTS[act](conv_circular(h, w, b))
    return h


# =============================================================================
# coupling layers  (src/nn/scalar/couplings_.py)
# =============================================================================
def _sum_density(v):
    """Module_.sum_density, src/nn/_core.py:38-42."""
    return v.reshape(v.shape[0], -1).sum(axis=1)


def affine_atomic(x_active, out, mask, parity, log0, inverse=False):
    """AffineCoupling_.atomic_forward / atomic_backward, couplings_.py:123-139.
    `out` = conditioner output (B,2,*L)."""
    t = mask_purify(mask, out[:, 0], parity)
    s = mask_purify(mask, out[:, 1], parity)
    s = np.where(_re(s) >= 0, s, -s)
    if not inverse:
        return t + x_active * np.exp(-s), log0 - _sum_density(s)
    return (x_active - t) * np.exp(s), log0 + _sum_density(s)


def shift_atomic(x_active, out, mask, parity, log0, inverse=False):
    """ShiftCoupling_.atomic_forward / atomic_backward, couplings_.py:110-116."""
    t = out[:, 0]
    sign = -1.0 if inverse else 1.0
    return mask_purify(mask, x_active + sign * t, parity), log0


def rqs_atomic(x_active, out, mask, parity, log0, *, xlim, ylim, extrap,
               inverse=False, stable_inverse=False, knots_x=None, knots_y=None):
    """RQSplineCoupling_.atomic_forward / atomic_backward, couplings_.py:178-200."""
    kx, ky, kd = knots_from_raw(out, xlim=xlim, ylim=ylim, axis=1, knots_x=knots_x, knots_y=knots_y)
    spline = RQSpline(kx, ky, kd, axis=1, extrap=extrap, stable_inverse=stable_inverse)
    xa = x_active[:, None]
    fx, g = spline.backward(xa) if inverse else spline.forward(xa)
    fx, g = fx[:, 0], g[:, 0]
    fx = mask_purify(mask, fx, parity)
    with np.errstate(divide='ignore', invalid='ignore'):
        logg = np.log(g)
    # purify(log g) multiplies by 0 at frozen sites (couplings_.py:186-187)
    logg = np.where((mask if parity == 0 else 1 - mask) != 0, logg, 0.0)
    return fx, log0 + _sum_density(logg)


def coupling_forward(x, log0, mask, steps, inverse=False):
    """Coupling_.forward / backward, couplings_.py:54-78.

    `steps` is a list, one per atomic step, of callables
        step(x_active, x_frozen, parity, log0, inverse) -> (x_active_new, log0)
    Step k acts on partition k % 2; the inverse visits them in reverse.
    """
    parts = list(mask_split(mask, x))
    order = range(len(steps))
    for k in (reversed(order) if inverse else order):
        p = k % 2
        parts[p], log0 = steps[k](parts[p], parts[1 - p], p, log0, inverse)
    return parts[0] + parts[1], log0


def make_convact_step(kind, layers, acts, mask, **kw):
    """One atomic coupling step whose conditioner is a ConvAct stack."""
    def step(x_active, x_frozen, parity, log0, inverse):
        out = convact_forward(x_frozen[:, None], layers, acts)
        if kind == 'affine':
            return affine_atomic(x_active, out, mask, parity, log0, inverse)
        if kind == 'shift':
            return shift_atomic(x_active, out, mask, parity, log0, inverse)
        if kind == 'rqs':
            return rqs_atomic(x_active, out, mask, parity, log0, inverse=inverse, **kw)
        raise ValueError(kind)
    return step


def multi_rqs_atomic(x_active, out, mask, parity, log0, *, xlims, ylims, extraps, inverse=False):
    """MultiRQSplineCoupling_.atomic_forward / atomic_backward, couplings_.py:313-336 with
    make_spline :349-413 (no fixed knots): data (B,S,*L); `out` (B, S*(3K-2), *L) split into S equal
    blocks, block i parametrising the spline of component i; log g summed over components."""
    S = len(xlims)
    P = out.shape[1] // S
    ys = []
    for i in range(S):
        yi, log0 = rqs_atomic(x_active[:, i], out[:, i * P:(i + 1) * P], mask, parity, log0,
                              xlim=xlims[i], ylim=ylims[i], extrap=extraps[i], inverse=inverse)
        ys.append(yi)
    return np.stack(ys, axis=1), log0


def make_multi_rqs_step(layers, acts, mask, **kw):
    """Atomic step of MultiRQSplineCoupling_ whose conditioner takes the S components of the frozen
    field as input channels (the (B,1,S,*L) tensor of preprocess_fz with its unit axis squeezed)."""
    def step(x_active, x_frozen, parity, log0, inverse):
        out = convact_forward(x_frozen, layers, acts)
        return multi_rqs_atomic(x_active, out, mask, parity, log0, inverse=inverse, **kw)
    return step


def cntr_coupling_forward(x, control, log0, mask, steps, inverse=False):
    """DirectCntrCoupling_.forward / backward, cntr_couplings_.py:20-52: as coupling_forward, but
    step 0 is conditioned on `control` instead of the frozen partition."""
    parts = list(mask_split(mask, x))
    order = range(len(steps))
    for k in (reversed(order) if inverse else order):
        p = k % 2
        frozen = control if k == 0 else parts[1 - p]
        parts[p], log0 = steps[k](parts[p], frozen, p, log0, inverse)
    return parts[0] + parts[1], log0


# =============================================================================
# DistConvertor_  (src/nn/scalar/modules_.py:93-114, 277-302, 333-383)
# =============================================================================
def expit_forward(x, log0):
    """Expit_.forward, modules_.py:96-99."""
    y = 1 / (1 + np.exp(-x))
    return y, log0 + _sum_density(-x + 2 * np.log(y))


def logit_forward(x, log0):
    """Logit_.forward, modules_.py:108-111."""
    y = np.log(x / (1 - x))
    return y, log0 - _sum_density(np.log(x * (1 - x)))


def splinenet_knots(weights_x, weights_y, weights_d, *, xlim, ylim):
    """SplineNet.make_spline, src/nn/scalar/modules.py:369-391 for
    spline_shape=[]: softmax over dim 0; weights_d None -> smooth derivatives."""
    kx = to_coord(weights_x, 0) * (xlim[1] - xlim[0]) + xlim[0]
    ky = to_coord(weights_y, 0) * (ylim[1] - ylim[0]) + ylim[0]
    kd = None if weights_d is None else softplus_ln2(weights_d)
    return kx, ky, kd


def splinenet_forward(x, log0, knots, extrap, inverse=False, stable_inverse=False):
    """SplineNet_.forward / backward, modules_.py:284-302 (spline on x.ravel())."""
    kx, ky, kd = knots
    spline = RQSpline(kx, ky, kd, extrap=extrap, stable_inverse=stable_inverse)
    fx, g = spline.backward(x.ravel()) if inverse else spline.forward(x.ravel())
    fx, g = fx.reshape(x.shape), g.reshape(x.shape)
    return fx, log0 + _sum_density(np.log(g))


def distconvertor(x, log0, weights, *, symmetric, inverse=False, stable_inverse=False):
    """DistConvertor_ = [Expit_, SplineNet_, Logit_], modules_.py:333-361
    (sgnbias/scale options off).  `weights` = (weights_x, weights_y, weights_d|None)."""
    if symmetric:
        lim, extrap = (0.5, 1.0), {'left': 'anti'}
    else:
        lim, extrap = (0.0, 1.0), {}
    knots = splinenet_knots(*weights, xlim=lim, ylim=lim)
    # forward: expit -> spline -> logit ; backward (ModuleList_.backward,
    # nn/_core.py:69-72) visits them reversed, each `.backward`:
    #   Logit_.backward = Expit_.forward, SplineNet_.backward, Expit_.backward = Logit_.forward
    x, log0 = expit_forward(x, log0)
    x, log0 = splinenet_forward(x, log0, knots, extrap, inverse=inverse,
                                stable_inverse=stable_inverse)
    return logit_forward(x, log0)


# =============================================================================
# PSDBlock_ = MeanFieldNet_ + FFTNet_  (src/nn/scalar/psd_.py, meanfield_.py, fftflow_.py)
# =============================================================================
def lattice_k2(lat_shape):
    """FreeScalar.calc_lattice_k2 + outer_lattice_k2, fftflow_.py:317-349: sum over axes of
    4 sin^2(k/2) with k = linspace(0, 2 pi (1 - 1/n), n), last axis trimmed to n//2 + 1."""
    grids = [4 * np.sin(np.linspace(0, 2 * np.pi * (1 - 1 / n), n) / 2) ** 2 for n in lat_shape]
    total = grids[0]
    for g in grids[1:]:
        total = total[..., None] + g
    return total[..., :(1 + lat_shape[-1] // 2)]


def ipsd_forward(norm_k2, weights, logy, ignore_zeromode=False):
    """IPSD.forward, fftflow_.py:232-239: y0 + y1 * SplineNet(norm_k2) with SplineNet's default
    limits (0, 1) and no extrapolation (modules.py:318-330); the k = 0 entry set to 1 when
    the zero mode is ignored.  `weights` = (weights_x, weights_y, weights_d | None)."""
    knots = splinenet_knots(*weights, xlim=(0.0, 1.0), ylim=(0.0, 1.0))
    spline = RQSpline(*knots, extrap={})
    s, _ = spline.forward(norm_k2.ravel())
    sigma = np.exp(logy[0]) + np.exp(logy[1]) * s.reshape(norm_k2.shape)
    if ignore_zeromode:
        sigma[tuple([0] * norm_k2.ndim)] = 1
    return sigma


def fftnet_log_jacobian(w):
    """FFTNet_.log_jacobian, fftflow_.py:167-180 (all axes of w are lattice axes)."""
    return 2 * np.sum(np.log(w)) - np.sum(np.log(w[..., 0:1])) - np.sum(np.log(w[..., -1:]))


def fftnet(x, log0, ipsd, inverse=False):
    """FFTNet_.forward / backward, fftflow_.py:121-131."""
    w = 1 / ipsd ** 0.5
    axes = tuple(range(-ipsd.ndim, 0))
    spec = np.fft.rfftn(x, axes=axes)
    spec = spec / w if inverse else spec * w
    logj = fftnet_log_jacobian(w)
    return np.fft.irfftn(spec, axes=axes), (log0 - logj if inverse else log0 + logj)


def scalenet(x, log0, raw_weight, inverse=False):
    """ScaleNet_, modules_.py:44-69."""
    w = softplus_ln2(raw_weight)
    nvar = np.prod(x.shape[1:])
    if inverse:
        return x / w, log0 - np.log(w) * nvar * np.ones(x.shape[0])
    return x * w, log0 + np.log(w) * nvar * np.ones(x.shape[0])


def meanfieldnet(x_mean, log0, rvol, weights, *, symmetric, final_scale=None, inverse=False):
    """MeanFieldNet_.forward / backward with `rvol` given (meanfield_.py:33-36, 49-52): the
    DistConvertor_ [Expit_, SplineNet_, Logit_ (, ScaleNet_)] acts on x_mean * rvol.
    `final_scale`: raw `_weight` of the trailing ScaleNet_ or None."""
    v = x_mean * rvol
    if inverse:
        if final_scale is not None:
            v, log0 = scalenet(v, log0, final_scale, inverse=True)
        v, log0 = distconvertor(v, log0, weights, symmetric=symmetric, inverse=True)
    else:
        v, log0 = distconvertor(v, log0, weights, symmetric=symmetric)
        if final_scale is not None:
            v, log0 = scalenet(v, log0, final_scale)
    return v / rvol, log0


def psdblock(x, log0, *, mf_weights, mf_symmetric, mf_final_scale, ipsd, inverse=False):
    """PSDBlock_.forward / backward, psd_.py:25-40.  mf_weights=None stands for Identity_."""
    axes = tuple(range(1, x.ndim))
    rvol = np.prod(x.shape[1:]) ** 0.5
    x_mean = np.mean(x, axis=axes).reshape(-1, *[1 for _ in axes])
    if mf_weights is None:
        y_mf, logj_mf = x_mean, 0
    else:
        y_mf, logj_mf = meanfieldnet(x_mean, 0, rvol, mf_weights, symmetric=mf_symmetric,
                                     final_scale=mf_final_scale, inverse=inverse)
    y_fft, logj_fft = fftnet(x - x_mean, 0, ipsd, inverse=inverse)
    return y_mf + y_fft, log0 + logj_mf + logj_fft


# =============================================================================
# Metropolis  (src/mcmc/mcmc.py)
# =============================================================================
def metropolis_accept_status(logqp, uniforms, logqp_ref=None):
    """Metropolis.calc_accept_status, src/mcmc/mcmc.py:304-317, with the
    `np.random.rand(B)` draw passed in explicitly as `uniforms`."""
    if logqp_ref is None:
        logqp_ref = logqp[0]
    status = np.empty(len(logqp), dtype=bool)
    lrand = np.log(uniforms)
    for i, l_i in enumerate(logqp):
        status[i] = lrand[i] < (logqp_ref - l_i)
        if status[i]:
            logqp_ref = l_i
    return status


def metropolis_accept_indices(accept_seq):
    """Metropolis.calc_accept_indices, src/mcmc/mcmc.py:319-328."""
    idx = np.arange(len(accept_seq))
    last = 0
    for i, acc in enumerate(accept_seq):
        if acc:
            last = i
        else:
            idx[i] = last
    return idx


def mcmc_accept_reject(y, logq, logp, uniforms, ref):
    """MCMCSampler._accept_reject_step, src/mcmc/mcmc.py:55-87.  `ref` is the
    chain state dict(sample, logq, logp, logqp) carried across calls (mutated)."""
    y, logq, logp = y.copy(), logq.copy(), logp.copy()
    acc = metropolis_accept_status(logq - logp, uniforms, ref.get('logqp'))
    if not acc[0]:
        y[0], logq[0], logp[0] = ref['sample'], ref['logq'], ref['logp']
    idx = metropolis_accept_indices(acc)
    y, logq, logp = y[idx], logq[idx], logp[idx]
    ref.update(sample=y[-1].copy(), logq=float(logq[-1]), logp=float(logp[-1]))
    ref['logqp'] = ref['logq'] - ref['logp']
    return y, logq, logp, acc, idx


# =============================================================================
# whole-path drivers
# =============================================================================
def blocked_mcmc(x, evaluate, proposals, log_uniforms, n_samples, n_blocks, logqp_ref=None):
    """BlockedMCMCSampler.sample__ + sweep, mcmc.py:143-219, for one chain.
    x: (1, *L) prior-space configuration (modified in place); evaluate(x) -> (y, logq, logp);
    proposals: iterator of block redraws (1, block_len); log_uniforms: iterator of arrays of
    n_blocks log-uniforms, one per sweep.  Returns cfgs, logq, logp, accept flags, last logqp_ref."""
    shape = x.shape
    block_len = x.size // n_blocks
    cfgs, logqs, logps = [], [], []
    accept = np.empty((n_samples, n_blocks), dtype=bool)
    for ind in range(n_samples):
        lrand = next(log_uniforms)
        for b in range(n_blocks):
            view = x.reshape(1, -1, block_len)
            backup = view[:, b].copy()
            view[:, b] = next(proposals)
            x = view.reshape(shape)
            _, logq, logp = evaluate(x)
            if b == 0 and logqp_ref is None:
                accept[ind, b] = True
            else:
                accept[ind, b] = lrand[b] < logqp_ref - (logq - logp)[0]
            if accept[ind, b]:
                logqp_ref = float((logq - logp)[0])
            else:
                view = x.reshape(1, -1, block_len)
                view[:, b] = backup
                x = view.reshape(shape)
        y, logq, logp = evaluate(x)
        cfgs.append(y[0])
        logqs.append(logq[0])
        logps.append(logp[0])
    return np.stack(cfgs), np.array(logqs), np.array(logps), accept, logqp_ref, x


def posterior_sample__(x, flow, action_kwargs, loc=0.0, scale=1.0):
    """Posterior.sample_ / sample__, src/_normflowcore.py:87-111, with the prior
    draw `x` passed in: returns (y, logq, logp)."""
    logr = normal_log_prob(x, loc, scale)
    y, logJ = flow(x, np.zeros(x.shape[0], dtype=logr.dtype))
    return y, logr - logJ, -phi4_action(y, **action_kwargs)


def kl_loss(logq, logp):
    """Fitter.calc_kl_mean, src/_normflowcore.py:325-329."""
    return np.mean(logq - logp)


def directional_derivative(fn, args, directions, h=1e-30):
    """Complex-step derivative d/de fn(*(a + e*v)) at e=0, exact to rounding for
    the analytic pieces (branches are taken on real parts).  Independent check of
    the hand-written backward kernels."""
    shifted = [np.asarray(a, dtype=np.complex128) + 1j * h * np.asarray(v)
               for a, v in zip(args, directions)]
    out = fn(*shifted)
    if isinstance(out, tuple):
        return tuple(np.imag(o) / h for o in out)
    return np.imag(out) / h
