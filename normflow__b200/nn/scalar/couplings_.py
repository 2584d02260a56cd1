"""Mask-based coupling layers (reference src/nn/scalar/couplings_.py).

The reference splits the field into two zero-interleaved halves, runs each atomic step
through ~10 (affine) to ~100 (spline) ATen launches and adds the halves back.  Here a
step is: conditioner -> ONE kernel that reads the field and the conditioner output,
tests the site's partition, applies the transform, passes frozen sites through and
reduces log|det J| per sample.  `Coupling_` itself keeps the reference's generic
split / atomic_* / cat dataflow for subclasses that only define `atomic_forward` and
`atomic_backward`.
"""

import os

import torch

from .._core import Module_
from ... import _ops, _C
from .modules import ConvAct
from ...mask import EvenOddMask


class Coupling_(Module_):
    """A list of atomic coupling steps acting alternately on the two partitions of
    `mask`; step k updates partition k % 2 conditioned on the other one
    (reference couplings_.py:22-103)."""

    def __init__(self, nets, *, mask, channels_axis=1, label='coupling_'):
        super().__init__(label=label)
        self.nets = torch.nn.ModuleList(nets)
        self.mask = mask
        self.channels_axis = channels_axis

    # ---- generic dataflow (reference couplings_.py:54-78) -------------------------
    def forward(self, x, log0=0):
        parts = list(self.mask.split(x))
        for k, net in enumerate(self.nets):
            p = k % 2
            parts[p], log0 = self.atomic_forward(x_active=parts[p], x_frozen=parts[1 - p],
                                                 parity=p, net=net, log0=log0)
        return self.mask.cat(*parts), log0

    def backward(self, x, log0=0):
        parts = list(self.mask.split(x))
        for k in reversed(range(len(self.nets))):
            p = k % 2
            parts[p], log0 = self.atomic_backward(x_active=parts[p], x_frozen=parts[1 - p],
                                                  parity=p, net=self.nets[k], log0=log0)
        return self.mask.cat(*parts), log0

    def atomic_forward(self, *, x_active, x_frozen, parity, net, log0=0):
        raise NotImplementedError

    def atomic_backward(self, *, x_active, x_frozen, parity, net, log0=0):
        raise NotImplementedError

    def preprocess_fz(self, x):
        return x.unsqueeze(self.channels_axis)

    def preprocess(self, x):
        return x.unsqueeze(self.channels_axis)

    def postprocess(self, x):
        return x.squeeze(self.channels_axis)

    def transfer(self, scale_factor=1, mask=None, **extra):
        return self.__class__([net.transfer(scale_factor=scale_factor) for net in self.nets],
                              mask=self.mask if mask is None else mask,
                              label=self.label, channels_axis=self.channels_axis)

    # ---- full-field fast path ------------------------------------------------------
    def _conditioner(self, net, x, parity):
        """Conditioner output for the step that updates partition `parity`: the net sees
        the frozen partition only, zero elsewhere, as one channel (couplings_.py:88-89).
        A ConvAct reads `x` directly and applies the partition test in its first layer."""
        if self.channels_axis != 1:
            raise NotImplementedError("the accelerated couplings assume channels_axis == 1")
        frozen_keep = 0 if parity == 0 else 1       # mask value of the frozen partition
        if isinstance(net, ConvAct) and net.fusable:
            return net.forward_masked(x, self.mask._mask, frozen_keep)
        return net(_ops.mask_select(x, self.mask._mask, frozen_keep).unsqueeze(1))

    def _sweep(self, x, log0, inverse):
        order = range(len(self.nets))
        for k in (reversed(order) if inverse else order):
            p = k % 2
            net = self.nets[k]
            if (not inverse and self._fused_kind is not None and torch.is_grad_enabled()
                    and self._fusable_train(net, x)):
                # training forward: the same fused kernel, which also keeps what the gradient kernels need
                convs = net._convs()
                x, log0 = _ops.fused2d_step_train(x, [c.weight for c in convs], [c.bias for c in convs],
                                                  self._fused_kind, self._fused_params(convs[-1].out_channels),
                                                  self.mask._mask, self.mask.mask_kwargs.get('parity', 0), p, log0)
                continue
            if (not inverse and self._fused_kind is not None and torch.is_grad_enabled()
                    and self._fusable_nd_train(net, x)):
                convs = net._convs()
                x, log0 = _ops.fusednd_step_train(x, [c.standard_weight() for c in convs], [c.bias for c in convs],
                                                  self._fused_kind, self._fused_params(convs[-1].out_channels),
                                                  self.mask._mask, self.mask.mask_kwargs.get('parity', 0), p, log0)
                continue
            if self._fused_kind is not None and self._fusable(net, x):
                # conditioner + transform in ONE kernel: the (B,P,*L) tensor is never formed
                convs = net._convs()
                x, log0 = _ops.fused2d_step(x, [c.weight for c in convs], [c.bias for c in convs],
                                            self._fused_kind, self._fused_params(convs[-1].out_channels),
                                            self.mask.mask_kwargs.get('parity', 0), p, log0, inverse)
                continue
            if self._fused_kind is not None and self._fusable_nd(net, x):
                convs = net._convs()
                x, log0 = _ops.fusednd_step(x, [c.standard_weight() for c in convs], [c.bias for c in convs],
                                            self._fused_kind, self._fused_params(convs[-1].out_channels),
                                            self.mask.mask_kwargs.get('parity', 0), p, log0, inverse)
                continue
            out = self._conditioner(net, x, p)
            x, log0 = self._transform(x, out, p, log0, _C.FROZEN_COPY, inverse)
        return x, log0

    _fused_kind = None          # 0 affine, 1 RQ spline; None: no fused kernel for this coupling

    def _fused_knots(self, net):
        return None             # number of spline knots the conditioner parametrises (None: affine)

    def _fused_params(self, n_channels):
        return None

    def _fusable(self, net, x):
        """The single-kernel step applies to evaluation (no autograd graph wanted) of a 2-D
        checkerboard coupling whose conditioner is ConvAct(1->8->8->P, 3x3, tanh)."""
        if not (isinstance(net, ConvAct) and net.fused2d_ok):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in net.parameters())):
            return False
        if not (isinstance(self.mask, EvenOddMask) and self.mask.mask_kwargs.get('exclude_mu') is None):
            return False
        return (x.dim() == 3 and x.is_cuda and self.channels_axis == 1
                and _ops.fused2d_supported(x.shape[1], x.shape[2], self._fused_knots(net)))

    def _fusable_nd(self, net, x):
        """Evaluation of a checkerboard coupling whose conditioner is ConvAct(1->H->H->P, 3^D taps, tanh) on a
        2-D .. 4-D lattice, H in {8, 16, 32, 64} (2-D with H = 8 has taken the single-kernel step before this is
        asked): layers 2 and 3 on the tensor cores, transform fused into the last one (nfk_fusednd_step).
        NFK_FUSED_ND=0 in the environment keeps the layer-by-layer kernels (A/B tests)."""
        if os.environ.get('NFK_FUSED_ND') == '0':
            return False
        if not (isinstance(net, ConvAct) and net.fusednd_ok):
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in net.parameters())):
            return False
        if not (isinstance(self.mask, EvenOddMask) and self.mask.mask_kwargs.get('exclude_mu') is None):
            return False
        return (x.is_cuda and self.channels_axis == 1 and x.dim() == net.conv_kwargs['conv_dim'] + 1
                and _ops.fusednd_supported(x.shape[1:], self._fused_knots(net), net.conv_kwargs['hidden_sizes'][0]))

    def _fusable_nd_train(self, net, x):
        """The same structural conditions with an autograd graph wanted: the N-D tensor-core forward that also keeps
        what the gradient kernels need (nfk_fusednd_step_train).  NFK_FUSED_ND_TRAIN=0 keeps the layer-by-layer path."""
        if os.environ.get('NFK_FUSED_ND_TRAIN') == '0' or os.environ.get('NFK_FUSED_ND') == '0':
            return False
        if not (isinstance(net, ConvAct) and net.fusednd_ok):
            return False
        if not (x.requires_grad or any(p.requires_grad for p in net.parameters())):
            return False
        if not (isinstance(self.mask, EvenOddMask) and self.mask.mask_kwargs.get('exclude_mu') is None):
            return False
        return (x.is_cuda and self.channels_axis == 1 and x.dim() == net.conv_kwargs['conv_dim'] + 1
                and _ops.fusednd_supported(x.shape[1:], self._fused_knots(net), net.conv_kwargs['hidden_sizes'][0]))

    def _fusable_train(self, net, x):
        """Same structural conditions, with an autograd graph wanted: needs the tensor-core kernel.
        NFK_FUSED_TRAIN=0 in the environment keeps the layer-by-layer kernels (A/B tests)."""
        if os.environ.get('NFK_FUSED_TRAIN') == '0':
            return False
        if not (isinstance(net, ConvAct) and net.fused2d_ok):
            return False
        if not (x.requires_grad or any(p.requires_grad for p in net.parameters())):
            return False
        if not (isinstance(self.mask, EvenOddMask) and self.mask.mask_kwargs.get('exclude_mu') is None):
            return False
        return (x.dim() == 3 and x.is_cuda and self.channels_axis == 1
                and _ops.fused2d_train_supported(x.shape[1], x.shape[2], self._fused_knots(net)))

    def _transform(self, x, out, parity, log0, frozen_mode, inverse):
        raise NotImplementedError


class ShiftCoupling_(Coupling_):
    """x_active +- t(x_frozen); unit Jacobian (reference couplings_.py:107-116)."""

    def forward(self, x, log0=0):
        return self._sweep(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._sweep(x, log0, inverse=True)

    def _transform(self, x, out, parity, log0, frozen_mode, inverse):
        return _ops.shift_apply(x, out, self.mask._mask, parity, frozen_mode, inverse), log0

    def atomic_forward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, False)

    def atomic_backward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, True)


class AffineCoupling_(Coupling_):
    """y = t + x exp(-|s|), log|J| = -sum |s|  (reference couplings_.py:120-139)."""

    _fused_kind = 0

    def forward(self, x, log0=0):
        return self._sweep(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._sweep(x, log0, inverse=True)

    def _transform(self, x, out, parity, log0, frozen_mode, inverse):
        return _ops.affine_apply(x, out, self.mask._mask, parity, log0, frozen_mode, inverse)

    def atomic_forward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, False)

    def atomic_backward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, True)


class RQSplineCoupling_(Coupling_):
    """Monotone rational-quadratic spline per site, knots from the conditioner
    (reference couplings_.py:143-275).  With K knots the conditioner emits 3K-2
    channels: K-1 bin widths and K-1 bin heights (softmax -> cumulative sum over
    `xlim` / `ylim`) and K knot derivatives (softplus with beta = ln 2, so a zero
    input gives derivative 1).

        extrap = {'left': 'linear', 'right': 'linear'}   # straight lines outside xlim
    """

    def __init__(self, nets, *, mask, xlim=(0, 1), ylim=(0, 1), knots_x=None, knots_y=None,
                 extrap={}, **kwargs):
        super().__init__(nets, mask=mask, **kwargs)
        self.xlim, self.xwidth = xlim, xlim[1] - xlim[0]
        self.ylim, self.ywidth = ylim, ylim[1] - ylim[0]
        self.knots_x, self.knots_y = knots_x, knots_y
        self.extrap = extrap
        # Fixed knots (couplings_.py:246-258): the conditioner then parametrises only the other coordinate and
        # the derivatives (2K-1 channels), or the derivatives alone (K channels).  The spline kernel takes bin
        # WIDTHS as softmax logits, so a fixed coordinate enters as the constant channels log(width_i) over the
        # limits (knots[0], knots[-1]) -- softmax of those logits gives back the widths to float32 rounding.
        self._fixed = {}
        for name, knots in (('x', knots_x), ('y', knots_y)):
            if knots is None:
                continue
            k = torch.as_tensor(knots, dtype=torch.float64).reshape(-1).cpu()
            if k.numel() < 2 or not bool((k[1:] > k[:-1]).all()):
                raise ValueError(f"knots_{name} must be a 1-D increasing sequence of at least 2 knots")
            logits = torch.log((k[1:] - k[:-1]) / (k[-1] - k[0])).to(torch.float32)
            self.register_buffer(f"_fixed_logits_{name}", logits, persistent=False)
            self._fixed[name] = (float(k[0]), float(k[-1]), k.numel())

    def forward(self, x, log0=0):
        return self._sweep(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._sweep(x, log0, inverse=True)

    _fused_kind = 1

    def _fused_params(self, n_channels):
        return _ops.rqs_params((n_channels + 2) // 3, self.xlim, self.ylim, self.extrap)

    def _fused_knots(self, net):
        if self._fixed:
            return -1                       # the single-kernel step parametrises all three knot arrays
        n = net.conv_kwargs['out_channels']
        return (n + 2) // 3 if (n + 2) % 3 == 0 else -1

    def _params(self, out):
        n = out.shape[self.channels_axis]
        if (n + 2) % 3 != 0:
            raise ValueError(f"conditioner emits {n} channels; an RQ spline needs 3K-2")
        xlim = self._fixed['x'][:2] if 'x' in self._fixed else self.xlim
        ylim = self._fixed['y'][:2] if 'y' in self._fixed else self.ylim
        return _ops.rqs_params((n + 2) // 3, xlim, ylim, self.extrap)

    def _with_fixed_knots(self, out):
        """Conditioner output -> the full (K-1, K-1, K) channel layout of the spline kernel."""
        if not self._fixed:
            return out
        ax = self.channels_axis
        n = out.shape[ax]
        both = len(self._fixed) == 2
        K = n if both else (n + 2) // 2
        if (not both and 2 * K - 1 != n) or any(v[2] != K for v in self._fixed.values()):
            raise ValueError(f"conditioner emits {n} channels, which does not match the fixed knots "
                             f"({ {k: v[2] for k, v in self._fixed.items()} } knots)")

        def const(name):
            shape = [1] * out.dim()
            shape[ax] = K - 1
            full = list(out.shape)
            full[ax] = K - 1
            return getattr(self, f"_fixed_logits_{name}").to(out.device).reshape(shape).expand(full)
        if both:
            parts = [const('x'), const('y'), out]
        else:
            free, d = out.split((K - 1, K), dim=ax)
            parts = [const('x'), free, d] if 'x' in self._fixed else [free, const('y'), d]
        return torch.cat(parts, dim=ax)

    def _transform(self, x, out, parity, log0, frozen_mode, inverse):
        out = self._with_fixed_knots(out)
        return _ops.rqs_apply(x, out, self.mask._mask, parity, self._params(out), log0, frozen_mode, inverse)

    def atomic_forward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, False)

    def atomic_backward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, True)

    def transfer(self, scale_factor=1, mask=None, **extra):
        return self.__class__([net.transfer(scale_factor=scale_factor) for net in self.nets],
                              mask=self.mask if mask is None else mask, label=self.label,
                              channels_axis=self.channels_axis, xlim=self.xlim, ylim=self.ylim,
                              knots_x=self.knots_x, knots_y=self.knots_y, extrap=self.extrap)


class MultiRQSplineCoupling_(Coupling_):
    """One rational-quadratic spline per extra channel of the data (reference
    couplings_.py:279-436): the field is (B, S, *L) with S = len(xlims) components, the
    conditioner sees the frozen partition of all components, (B, 1, S, *L), and emits
    S * (3K-2) channels, the i-th block of 3K-2 parametrising the spline of component i.

    The S components are S independent per-site splines, so they run as ONE launch of the
    spline kernel on the (B*S, *L) view of the field whenever the components share their limits
    and extrapolation (the default), and as S launches otherwise.
    """

    def __init__(self, nets, *, mask, xlims=[(0, 1), (0, 1)], ylims=[(0, 1), (0, 1)],
                 knots_x=[None, None], knots_y=[None, None], extraps=[{}, {}], **kwargs):
        super().__init__(nets, mask=mask, **kwargs)
        if any(k is not None for k in knots_x) or any(k is not None for k in knots_y):
            raise NotImplementedError("MultiRQSplineCoupling_ with fixed knots_x / knots_y is outside the "
                                      "accelerated hot path")
        if not (len(xlims) == len(ylims) == len(extraps)):
            raise ValueError("xlims, ylims and extraps must have one entry per spline")
        if self.channels_axis != 1:
            raise NotImplementedError("the accelerated couplings assume channels_axis == 1")
        self.num_splines = len(xlims)
        self.xlims, self.ylims = xlims, ylims
        self.xwidths = [lim[1] - lim[0] for lim in xlims]
        self.ywidths = [lim[1] - lim[0] for lim in ylims]
        self.knots_x, self.knots_y = knots_x, knots_y
        self.extraps = extraps

    def _uniform(self):
        first = (tuple(self.xlims[0]), tuple(self.ylims[0]), dict(self.extraps[0]))
        return all((tuple(a), tuple(b), dict(c)) == first
                   for a, b, c in zip(self.xlims, self.ylims, self.extraps))

    def _check(self, x, out):
        S = self.num_splines
        if x.dim() != 2 + len(self.mask.shape) or x.shape[1] != S:
            raise ValueError(f"MultiRQSplineCoupling_ with {S} splines expects data of shape (B, {S}, *lattice), "
                             f"got {tuple(x.shape)}")
        if out.shape[1] % S != 0 or (out.shape[1] // S + 2) % 3 != 0 or tuple(out.shape[2:]) != tuple(x.shape[2:]):
            raise ValueError(f"conditioner output {tuple(out.shape)}: need {S} x (3K-2) channels on the lattice")
        return out.shape[1] // S

    def _transform(self, x, out, parity, log0, frozen_mode, inverse):
        P = self._check(x, out)
        B, S, lat = x.shape[0], self.num_splines, tuple(x.shape[2:])
        K = (P + 2) // 3
        if self._uniform():
            prm = _ops.rqs_params(K, self.xlims[0], self.ylims[0], self.extraps[0])
            y, logj = _ops.rqs_apply(x.reshape(B * S, *lat), out.reshape(B * S, P, *lat), self.mask._mask,
                                     parity, prm, 0, frozen_mode, inverse)
            return y.reshape(B, S, *lat), log0 + logj.reshape(B, S).sum(dim=1)
        ys = []
        for i in range(S):
            prm = _ops.rqs_params(K, self.xlims[i], self.ylims[i], self.extraps[i])
            yi, log0 = _ops.rqs_apply(x[:, i].contiguous(), out[:, i * P:(i + 1) * P].contiguous(),
                                      self.mask._mask, parity, prm, log0, frozen_mode, inverse)
            ys.append(yi)
        return torch.stack(ys, dim=1), log0

    def _frozen(self, x, parity):
        B, S, lat = x.shape[0], x.shape[1], tuple(x.shape[2:])
        keep = 0 if parity == 0 else 1
        return _ops.mask_select(x.reshape(B * S, *lat), self.mask._mask, keep).reshape(B, S, *lat)

    def _multi_sweep(self, x, log0, inverse):
        order = range(len(self.nets))
        for k in (reversed(order) if inverse else order):
            p = k % 2
            out = self.nets[k](self.preprocess_fz(self._frozen(x, p)))
            x, log0 = self._transform(x, out, p, log0, _C.FROZEN_COPY, inverse)
        return x, log0

    def forward(self, x, log0=0):
        return self._multi_sweep(x, log0, inverse=False)

    def backward(self, x, log0=0):
        return self._multi_sweep(x, log0, inverse=True)

    def atomic_forward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, False)

    def atomic_backward(self, *, x_active, x_frozen, parity, net, log0=0):
        out = net(self.preprocess_fz(x_frozen))
        return self._transform(x_active, out, parity, log0, _C.FROZEN_ZERO, True)

    def preprocess(self, x):
        return torch.tensor_split(x, self.num_splines, dim=self.channels_axis)

    def postprocess(self, xs):
        return torch.cat(xs, dim=self.channels_axis)

    def transfer(self, scale_factor=1, mask=None, **extra):
        return self.__class__([net.transfer(scale_factor=scale_factor) for net in self.nets],
                              mask=self.mask if mask is None else mask, label=self.label,
                              channels_axis=self.channels_axis, xlims=self.xlims, ylims=self.ylims,
                              knots_x=self.knots_x, knots_y=self.knots_y, extraps=self.extraps)
