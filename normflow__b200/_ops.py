"""Thin torch.autograd wrappers over the C ABI (include/normflow_b200.h).

Every function here enqueues hand-written sm_100a kernels on torch's current CUDA
stream; tensors are only the memory they work on.  Nothing in this file computes on
the host and nothing falls back to eager PyTorch: a CPU tensor raises.
"""

import functools
import os
import numbers

import numpy as np
import torch

from . import _C
from ._C import dev, stream, check, lib


def _native(fn):
    """Run an op wrapper with __torch_function__ dispatch switched off.  Importing the package makes
    CUDA the default device (as the reference does), which installs a Python-level function mode
    that every torch call -- `.is_contiguous()`, `.data_ptr()`, `torch.empty` ... -- is routed
    through; the wrappers make a few dozen such calls per kernel launch and pass `device=`
    explicitly, so inside them the mode is pure overhead (it dominated the latency-bound configs)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with torch._C.DisableTorchFunction():
            # The C entry points launch on the CURRENT device and torch's current stream of it.  A model
            # that lives on another GPU (device_handler.to('cuda:1'), a snapshot loaded to cuda:{rank})
            # is served by making its device current for the call, as ATen does for its own kernels.
            index = _cuda_index(args, kwargs)
            if index >= 0 and index != torch._C._cuda_getDevice():
                with torch.cuda.device(index):
                    return fn(*args, **kwargs)
            return fn(*args, **kwargs)
    return wrapper


def _cuda_index(args, kwargs):
    """Device index of the first CUDA tensor (or explicit CUDA torch.device) among the arguments, -1 if none."""
    for a in args:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.get_device()
        elif isinstance(a, torch.device) and a.type == 'cuda' and a.index is not None:
            return a.index
    if kwargs:
        return _cuda_index(tuple(kwargs.values()), None)
    return -1


def _f32c(t, name):
    """Contiguous float32 view of a CUDA tensor (copying only when it has to)."""
    if not torch.is_tensor(t):
        raise TypeError(f"normflow_b200: {name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"normflow_b200: {name} is on '{t.device}'; the hot path runs on CUDA only "
                           "(no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"normflow_b200: {name} must be float32 (got {t.dtype}); the kernels are fp32")
    return t if t.is_contiguous() else t.contiguous()


def as_log(log0, like):
    """The reference threads `log0=0` (a python number) through the flow; the kernels
    take NULL for zero.  Returns a float32 [B] tensor or None."""
    if log0 is None:
        return None
    if isinstance(log0, numbers.Number):
        if log0 == 0:
            return None
        return torch.full((like.shape[0],), float(log0), dtype=torch.float32, device=like.device)
    return _f32c(log0, "log0")


def _mask_u8(mask):
    if mask.dtype != torch.uint8:
        raise TypeError("normflow_b200: mask must be uint8 (Mask._mask)")
    return mask if mask.is_contiguous() else mask.contiguous()


def _no_grad_needed(*tensors):
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "normflow_b200: the inverse direction is a sampling/evaluation path and has no "
            "backward kernel; call it under torch.no_grad()")


# ---------------------------------------------------------------------------- masks
class _MaskSelect(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, mask, keep):
        x = _f32c(x, "x")
        y = torch.empty_like(x)
        # the mask covers the trailing (lattice) axes; every leading axis -- batch, channels -- is a row
        V = mask.numel()
        if x.ndim < mask.ndim or tuple(x.shape[x.ndim - mask.ndim:]) != tuple(mask.shape):
            raise ValueError(f"field of shape {tuple(x.shape)} does not end with the mask's shape {tuple(mask.shape)}")
        check(lib().nfk_mask_select(dev(x), dev(mask, torch.uint8), keep, dev(y), x.numel() // max(V, 1), V,
                                    stream()), "mask_select")
        ctx.mask, ctx.keep = mask, keep
        return y

    @staticmethod
    @_native
    def backward(ctx, gy):
        return _MaskSelect.apply(gy, ctx.mask, ctx.keep), None, None


@_native
def mask_select(x, mask, keep):
    """Mask.split / purify: x where mask == keep, 0 elsewhere (mask/mask.py:30-37)."""
    return _MaskSelect.apply(x, _mask_u8(mask), int(keep))


@_native
def make_evenodd_mask(shape, parity, exclude_mu, device):
    m = torch.empty(tuple(shape), dtype=torch.uint8, device=device)
    check(lib().nfk_mask_evenodd(dev(m, torch.uint8), _C.lattice(shape), int(parity),
                                 -1 if exclude_mu is None else int(exclude_mu), stream()), "mask_evenodd")
    return m


@_native
def make_alongaxis_mask(shape, parity, mu, device):
    m = torch.empty(tuple(shape), dtype=torch.uint8, device=device)
    check(lib().nfk_mask_alongaxis(dev(m, torch.uint8), _C.lattice(shape), int(parity), int(mu), stream()),
          "mask_alongaxis")
    return m


# ---------------------------------------------------------------------------- prior
@_native
def prior_sample(batch_size, shape, loc, scale, seed, offset, device, with_logprob=True, state=None):
    """x = loc + scale * N(0,1) and (optionally) its log-density summed per sample.
    `state`: int64[2] CUDA tensor {seed, offset} -- the generator state kept on the device and
    advanced by the draw (CUDA-graph capturable); otherwise `seed` / `offset` are host integers."""
    shape = tuple(int(v) for v in shape)
    V = int(np.prod(shape)) if len(shape) else 1
    x = torch.empty((batch_size,) + shape, dtype=torch.float32, device=device)
    logr = torch.empty((batch_size,), dtype=torch.float32, device=device) if with_logprob else None
    if state is not None:
        with _C.timed("prior_normal_sample"):
            check(lib().nfk_prior_normal_sample_dev(dev(x), dev(logr), batch_size, V, dev(loc), dev(scale),
                                                    dev(state, torch.int64), stream()), "prior_normal_sample_dev")
        return x, logr
    with _C.timed("prior_normal_sample"):
        check(lib().nfk_prior_normal_sample(dev(x), dev(logr), batch_size, V, dev(loc), dev(scale),
                                            int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), stream()),
              "prior_normal_sample")
    return x, logr


@_native
def prior_logprob(x, loc, scale):
    x = _f32c(x, "x")
    B = x.shape[0]
    logr = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().nfk_prior_normal_logprob(dev(x), dev(logr), B, x.numel() // max(B, 1), dev(loc), dev(scale),
                                         stream()), "prior_normal_logprob")
    return logr


# ---------------------------------------------------------------------------- couplings
def _bv(x):
    B = x.shape[0]
    return B, x.numel() // max(B, 1)


class _AffineFwd(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, out, log_in, mask, parity, frozen_mode):
        x, out = _f32c(x, "x"), _f32c(out, "conditioner output")
        B, V = _bv(x)
        if out.numel() != 2 * B * V:
            raise ValueError(f"affine coupling needs 2 conditioner channels, got shape {tuple(out.shape)}")
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        check(lib().nfk_affine_fwd(dev(x), dev(out), dev(mask, torch.uint8), parity, frozen_mode, dev(log_in),
                                   dev(y), dev(log_out), B, V, stream()), "affine_fwd")
        ctx.save_for_backward(x, out, mask)
        ctx.cfg = (parity, frozen_mode, log_in is not None)
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        x, out, mask = ctx.saved_tensors
        parity, frozen_mode, has_log = ctx.cfg
        B, V = _bv(x)
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx, gout = torch.empty_like(x), torch.empty_like(out)
        check(lib().nfk_affine_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, frozen_mode, dev(gy),
                                   dev(glog), dev(gx), dev(gout), B, V, stream()), "affine_bwd")
        return gx, gout, (glog if has_log else None), None, None, None


@_native
def affine_apply(x, out, mask, parity, log0=0, frozen_mode=_C.FROZEN_COPY, inverse=False):
    """AffineCoupling_.atomic_forward/backward (couplings_.py:123-139)."""
    mask = _mask_u8(mask)
    log_in = as_log(log0, x)
    if not inverse:
        return _AffineFwd.apply(x, out, log_in, mask, int(parity), int(frozen_mode))
    _no_grad_needed(x, out, log_in)
    x, out = _f32c(x, "x"), _f32c(out, "conditioner output")
    B, V = _bv(x)
    y = torch.empty_like(x)
    log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().nfk_affine_inv(dev(x), dev(out), dev(mask, torch.uint8), int(parity), int(frozen_mode),
                               dev(log_in), dev(y), dev(log_out), B, V, stream()), "affine_inv")
    return y, log_out


class _Shift(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, out, mask, parity, frozen_mode, sign):
        x, out = _f32c(x, "x"), _f32c(out, "conditioner output")
        B, V = _bv(x)
        if out.numel() != B * V:
            raise ValueError(f"shift coupling needs 1 conditioner channel, got shape {tuple(out.shape)}")
        y = torch.empty_like(x)
        check(lib().nfk_shift_apply(dev(x), dev(out), dev(mask, torch.uint8), parity, frozen_mode, sign, dev(y),
                                    B, V, stream()), "shift_apply")
        ctx.mask = mask
        ctx.cfg = (parity, frozen_mode, sign, out.shape)
        return y

    @staticmethod
    @_native
    def backward(ctx, gy):
        parity, frozen_mode, sign, oshape = ctx.cfg
        gy = _f32c(gy, "gy")
        active = 1 if parity == 0 else 0
        g_act = mask_select(gy, ctx.mask, active)
        gx = gy if frozen_mode == _C.FROZEN_COPY else g_act
        return gx, (sign * g_act).reshape(oshape), None, None, None, None


@_native
def shift_apply(x, out, mask, parity, frozen_mode=_C.FROZEN_COPY, inverse=False):
    """ShiftCoupling_.atomic_forward/backward (couplings_.py:110-116)."""
    return _Shift.apply(x, out, _mask_u8(mask), int(parity), int(frozen_mode), -1.0 if inverse else 1.0)


def rqs_params(n_knots, xlim, ylim, extrap):
    extrap = extrap or {}
    for side in ('left', 'right'):
        if extrap.get(side) not in (None, 'linear'):
            raise NotImplementedError(
                f"RQSplineCoupling_: extrap[{side!r}]={extrap.get(side)!r} is not available in the fused "
                "coupling kernel (supported: None, 'linear')")
    return _C.RqsParams(int(n_knots), float(xlim[0]), float(xlim[1]), float(ylim[0]), float(ylim[1]),
                        _C.EXTRAP[extrap.get('left')], _C.EXTRAP[extrap.get('right')])


class _RqsFwd(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, out, log_in, mask, parity, frozen_mode, prm):
        x, out = _f32c(x, "x"), _f32c(out, "conditioner output")
        B, V = _bv(x)
        if out.numel() != (3 * prm.n_knots - 2) * B * V:
            raise ValueError(f"RQ-spline coupling with {prm.n_knots} knots needs {3 * prm.n_knots - 2} "
                             f"conditioner channels, got shape {tuple(out.shape)}")
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        with _C.timed("rqs_fwd"):
            check(lib().nfk_rqs_fwd(dev(x), dev(out), dev(mask, torch.uint8), parity, frozen_mode, prm,
                                    dev(log_in), dev(y), dev(log_out), B, V, stream()), "rqs_fwd")
        ctx.save_for_backward(x, out, mask)
        ctx.cfg = (parity, frozen_mode, prm, log_in is not None)
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        x, out, mask = ctx.saved_tensors
        parity, frozen_mode, prm, has_log = ctx.cfg
        B, V = _bv(x)
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx, gout = torch.empty_like(x), torch.empty_like(out)
        with _C.timed("rqs_bwd"):
            check(lib().nfk_rqs_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, frozen_mode, prm, dev(gy),
                                    dev(glog), dev(gx), dev(gout), B, V, stream()), "rqs_bwd")
        return gx, gout, (glog if has_log else None), None, None, None, None


@_native
def rqs_apply(x, out, mask, parity, prm, log0=0, frozen_mode=_C.FROZEN_COPY, inverse=False):
    """RQSplineCoupling_.atomic_forward/backward incl. make_spline (couplings_.py:178-262)."""
    mask = _mask_u8(mask)
    log_in = as_log(log0, x)
    if not inverse:
        return _RqsFwd.apply(x, out, log_in, mask, int(parity), int(frozen_mode), prm)
    _no_grad_needed(x, out, log_in)
    x, out = _f32c(x, "x"), _f32c(out, "conditioner output")
    B, V = _bv(x)
    y = torch.empty_like(x)
    log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().nfk_rqs_inv(dev(x), dev(out), dev(mask, torch.uint8), int(parity), int(frozen_mode), prm,
                            dev(log_in), dev(y), dev(log_out), B, V, stream()), "rqs_inv")
    return y, log_out


# ---------------------------------------------------------------------------- pointwise chains
class _Logistic(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, log_in, which):
        x = _f32c(x, "x")
        B, V = _bv(x)
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        check(lib().nfk_logistic_fwd(dev(x), which, dev(log_in), dev(y), dev(log_out), B, V, stream()),
              "logistic_fwd")
        ctx.save_for_backward(x)
        ctx.cfg = (which, log_in is not None)
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        (x,) = ctx.saved_tensors
        which, has_log = ctx.cfg
        B, V = _bv(x)
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx = torch.empty_like(x)
        check(lib().nfk_logistic_bwd(dev(x), which, dev(gy), dev(glog), dev(gx), B, V, stream()), "logistic_bwd")
        return gx, (glog if has_log else None), None


@_native
def logistic(x, which, log0=0):
    """Expit_ (which=0) / Logit_ (which=1) forward with log-Jacobian (modules_.py:93-114)."""
    return _Logistic.apply(x, as_log(log0, x), int(which))


class _Spline1d(torch.autograd.Function):
    """knots: float32 [3, K] (kx | ky | kd) or, for the logistic chain, [5, K] with the
    complements cx = x_hi - kx and cy = y_hi - ky appended."""

    @staticmethod
    @_native
    def forward(ctx, x, knots, log_in, left, right, logistic_wrap):
        x, knots = _f32c(x, "x"), _f32c(knots, "knots")
        K = knots.shape[1]
        B, V = _bv(x)
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        check(lib().nfk_spline1d_fwd(dev(x), dev(knots), K, left, right, logistic_wrap, 0,
                                     dev(log_in), dev(y), dev(log_out), B, V, stream()), "spline1d_fwd")
        ctx.save_for_backward(x, knots)
        ctx.cfg = (left, right, logistic_wrap, log_in is not None)
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        x, knots = ctx.saved_tensors
        left, right, logistic_wrap, has_log = ctx.cfg
        K = knots.shape[1]
        B, V = _bv(x)
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx = torch.empty_like(x)
        gk = torch.zeros((5, K), dtype=torch.float32, device=x.device)
        check(lib().nfk_spline1d_bwd(dev(x), dev(knots), K, left, right, logistic_wrap,
                                     dev(gy), dev(glog), dev(gx), dev(gk), B, V, stream()), "spline1d_bwd")
        return gx, gk[:knots.shape[0]], (glog if has_log else None), None, None, None


def _check_knots(knots, logistic_wrap):
    rows = 5 if logistic_wrap else 3
    if knots.dim() != 2 or knots.shape[0] != rows:
        raise ValueError(f"spline1d expects knots of shape [{rows}, K], got {tuple(knots.shape)}")
    if not 2 <= knots.shape[1] <= 64:
        raise ValueError("spline1d supports 2..64 knots")


@_native
def spline1d(x, knots, log0=0, extrap=None, logistic_wrap=False, inverse=False):
    """One shared 1-D RQ spline over every element (SplineNet_), optionally wrapped as
    Expit_ -> spline -> Logit_ (DistConvertor_) in a single kernel.  `knots` is the
    stacked [3, K] (or [5, K] for the chain) tensor described in _Spline1d."""
    extrap = extrap or {}
    left, right = _C.EXTRAP[extrap.get('left')], _C.EXTRAP[extrap.get('right')]
    log_in = as_log(log0, x)
    _check_knots(knots, logistic_wrap)
    if not inverse:
        return _Spline1d.apply(x, knots, log_in, left, right, int(bool(logistic_wrap)))
    _no_grad_needed(x, knots, log_in)
    x, knots = _f32c(x, "x"), _f32c(knots, "knots")
    B, V = _bv(x)
    y = torch.empty_like(x)
    log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().nfk_spline1d_fwd(dev(x), dev(knots), knots.shape[1], left, right, int(bool(logistic_wrap)), 1,
                                 dev(log_in), dev(y), dev(log_out), B, V, stream()), "spline1d_fwd(inverse)")
    return y, log_out


# ---------------------------------------------------------------------------- action
class _Phi4(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, phi, w0, w2, w4):
        phi = _f32c(phi, "cfgs")
        B = phi.shape[0]
        lat = _C.lattice(phi.shape[1:])
        S = torch.empty((B,), dtype=torch.float32, device=phi.device)
        with _C.timed("phi4_action_fwd"):
            check(lib().nfk_phi4_action_fwd(dev(phi), lat, w0, w2, w4, dev(S), B, stream()), "phi4_action_fwd")
        ctx.save_for_backward(phi)
        ctx.cfg = (lat, w0, w2, w4)
        return S

    @staticmethod
    @_native
    def backward(ctx, gS):
        (phi,) = ctx.saved_tensors
        lat, w0, w2, w4 = ctx.cfg
        gS = _f32c(gS, "gS")
        gphi = torch.empty_like(phi)
        check(lib().nfk_phi4_action_bwd(dev(phi), lat, w0, w2, w4, dev(gS), dev(gphi), phi.shape[0], stream()),
              "phi4_action_bwd")
        return gphi, None, None, None


@_native
def phi4_action(cfgs, w0, w2, w4):
    """ScalarPhi4Action.action (scalar_action.py:38-46)."""
    if cfgs.ndim < 2:
        raise ValueError("cfgs must have a batch axis and at least one lattice axis")
    return _Phi4.apply(cfgs, float(w0), float(w2), float(w4))


# ---------------------------------------------------------------------------- knot table
class _KnotTable(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, wx, wy, wd, lims):
        wx, wy = _f32c(wx, "weights_x"), _f32c(wy, "weights_y")
        wd = None if wd is None else _f32c(wd, "weights_d")
        K = wx.numel() + 1
        if wy.numel() != K - 1 or (wd is not None and wd.numel() != K):
            raise ValueError("weights_x / weights_y need K-1 entries and weights_d K")
        table = torch.empty((5, K), dtype=torch.float32, device=wx.device)
        check(lib().nfk_knots_fwd(dev(wx), dev(wy), dev(wd), K, *lims, dev(table), stream()), "knots_fwd")
        ctx.save_for_backward(wx, wy, wd)
        ctx.cfg = (K, lims)
        return table

    @staticmethod
    @_native
    def backward(ctx, g):
        wx, wy, wd = ctx.saved_tensors
        K, lims = ctx.cfg
        g = _f32c(g, "g_table")
        gwx, gwy = torch.empty_like(wx), torch.empty_like(wy)
        gwd = None if wd is None else torch.empty_like(wd)
        check(lib().nfk_knots_bwd(dev(wx), dev(wy), dev(wd), K, *lims, dev(g), dev(gwx), dev(gwy), dev(gwd),
                                  stream()), "knots_bwd")
        return gwx, gwy, gwd, None


@_native
def knot_table(weights_x, weights_y, weights_d, xlim, ylim):
    """SplineNet.make_spline for one shared spline (modules.py:369-391) in one launch: float32[5, K]
    = knots_x | knots_y | knots_d | xlim[1] - knots_x | ylim[1] - knots_y (the last two summed from
    the right end).  weights_d=None -> smooth derivatives."""
    lims = (float(xlim[0]), float(xlim[1] - xlim[0]), float(ylim[0]), float(ylim[1] - ylim[0]))
    return _KnotTable.apply(weights_x, weights_y, weights_d, lims)


# ---------------------------------------------------------------------------- PSD block
class _PsdWeights(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, ipsd, last_half, inverse):
        ipsd = _f32c(ipsd, "ipsd")
        Kc = ipsd.numel()
        w = torch.empty_like(ipsd)
        logj = torch.empty((1,), dtype=torch.float32, device=ipsd.device)
        check(lib().nfk_psd_weights_fwd(dev(ipsd), Kc, last_half, inverse, dev(w), dev(logj), stream()),
              "psd_weights_fwd")
        ctx.save_for_backward(ipsd, w)
        ctx.cfg = (Kc, last_half, inverse)
        return w, logj

    @staticmethod
    @_native
    def backward(ctx, gw, glogj):
        ipsd, w = ctx.saved_tensors
        Kc, last_half, inverse = ctx.cfg
        gw = None if gw is None else _f32c(gw, "gw")
        glogj = None if glogj is None else _f32c(glogj, "glogj")
        g = torch.empty_like(ipsd)
        check(lib().nfk_psd_weights_bwd(dev(ipsd), dev(w), dev(gw), dev(glogj), Kc, last_half, inverse,
                                        dev(g), stream()), "psd_weights_bwd")
        return g, None, None


@_native
def psd_weights(ipsd, inverse=False):
    """Spectral weights w = ipsd^(-1/2) (inverse: ipsd^(+1/2)) with the shape of `ipsd`
    (the rfftn half-spectrum of the lattice) and the 0-dim log-Jacobian of multiplying the
    spectrum by them (FFTNet_.forward/backward + log_jacobian, fftflow_.py:121-131,167-178)."""
    if ipsd.ndim < 1:
        raise ValueError("ipsd must have the shape of the half-spectrum")
    w, logj = _PsdWeights.apply(ipsd, int(ipsd.shape[-1]), 1 if inverse else 0)
    return w, logj.reshape(())


def _c64(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"normflow_b200: {name} is on '{t.device}'; the hot path runs on CUDA only "
                           "(no CPU fallback)")
    if t.dtype != torch.complex64:
        raise TypeError(f"normflow_b200: {name} must be complex64 (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


class _PsdScale(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, X, w, zero_mode, zero_scale, inplace):
        X = _c64(X, "spectrum")
        w = _f32c(w, "w")
        Kc = w.numel()
        B = X.numel() // Kc
        if zero_mode is not None:
            zero_mode = _f32c(zero_mode, "zero_mode").reshape(-1)
            if zero_mode.numel() != B:
                raise ValueError("zero_mode must hold one value per sample")
        Y = X if inplace else torch.empty_like(X)
        with _C.timed("psd_scale"):
            check(lib().nfk_psd_scale(dev(X, torch.complex64), dev(w), dev(zero_mode), zero_scale,
                                      dev(Y, torch.complex64), B, Kc, stream()), "psd_scale")
        if inplace:
            ctx.mark_dirty(X)
        ctx.save_for_backward(X if ctx.needs_input_grad[1] else None, w)
        ctx.cfg = (B, Kc, zero_mode is not None, zero_scale)
        return Y

    @staticmethod
    @_native
    def backward(ctx, gY):
        X, w = ctx.saved_tensors
        B, Kc, replace, zero_scale = ctx.cfg
        gY = _c64(gY, "gY")
        need_x, need_w, need_z = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gX = torch.empty_like(gY) if need_x else None
        gw = gw_part = g_zero = None
        if need_w:
            gw = torch.empty_like(w)
            gw_part = torch.empty((int(lib().nfk_psd_chunks(B, Kc)), Kc), dtype=torch.float32, device=w.device)
        if replace:
            g_zero = torch.empty((B,), dtype=torch.float32, device=w.device)
        check(lib().nfk_psd_scale_bwd(dev(X, torch.complex64) if need_w else None, dev(gY, torch.complex64),
                                      dev(w), 1 if replace else 0, zero_scale,
                                      dev(gX, torch.complex64) if need_x else None, dev(gw), dev(gw_part),
                                      dev(g_zero), B, Kc, stream()), "psd_scale_bwd")
        return gX, gw, (g_zero if need_z else None), None, None


@_native
def psd_scale(X, w, zero_mode=None, zero_scale=1.0):
    """Y = X * w over the trailing (half-spectrum) axes of the complex64 tensor X; with
    `zero_mode` (one float per sample) the k = 0 element of every sample is replaced by
    zero_scale * zero_mode.  Works in place when no gradient is being recorded."""
    if tuple(X.shape[X.ndim - w.ndim:]) != tuple(w.shape):
        raise ValueError(f"spectrum {tuple(X.shape)} does not end with the weights' shape {tuple(w.shape)}")
    recording = torch.is_grad_enabled() and (X.requires_grad or w.requires_grad or
                                             (zero_mode is not None and zero_mode.requires_grad))
    return _PsdScale.apply(X, w, zero_mode, float(zero_scale), not recording and X.is_contiguous())


def _sample_sum(x, scale):
    B = x.shape[0]
    V = x.numel() // max(B, 1)
    out = torch.empty((B,), dtype=torch.float32, device=x.device)
    check(lib().nfk_sample_mean(dev(x), B, V, scale, dev(out), stream()), "sample_mean")
    return out


class _SampleMean(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x):
        x = _f32c(x, "x")
        ctx.shape = tuple(x.shape)
        return _sample_sum(x, 1.0 / max(x.numel() // max(x.shape[0], 1), 1))

    @staticmethod
    @_native
    def backward(ctx, g):
        shape = ctx.shape
        V = int(np.prod(shape[1:]))
        return (g / V).reshape(-1, *[1] * (len(shape) - 1)).expand(shape)


class _SampleShift(torch.autograd.Function):
    @staticmethod
    @_native
    def forward(ctx, x, delta):
        x = _f32c(x, "x")
        delta = _f32c(delta, "delta").reshape(-1)
        B = x.shape[0]
        if delta.numel() != B:
            raise ValueError("delta must hold one value per sample")
        y = torch.empty_like(x)
        check(lib().nfk_sample_shift(dev(x), dev(delta), dev(y), B, x.numel() // max(B, 1), stream()),
              "sample_shift")
        ctx.dshape = None
        return y

    @staticmethod
    @_native
    def backward(ctx, gy):
        gy = _f32c(gy, "gy")
        gd = _sample_sum(gy, 1.0) if ctx.needs_input_grad[1] else None
        return gy, gd


@_native
def sample_mean(x):
    """Mean over everything but the batch axis -> float32[B] (meanfield_.py:29, psd_.py:28)."""
    return _SampleMean.apply(x)


@_native
def sample_shift(x, delta):
    """x[b, ...] + delta[b] (meanfield_.py:31)."""
    return _SampleShift.apply(x, delta)


# ---------------------------------------------------------------------------- conditioner
def _conv_call(inp, w, transposed, bias, in_mask, in_keep, act, dact_from, dact_kind, shape, ksize, Ci, Co,
               in_parity=None):
    B = inp.shape[0]
    out = torch.empty((B, Co) + tuple(shape), dtype=torch.float32, device=inp.device)
    if in_parity is not None and in_mask is None:
        # `inp` lives on one checkerboard partition only: products with the known zeros are skipped
        with _C.timed(f"conv_circ_fwd_cb[{Ci}->{Co}]"):
            check(lib().nfk_conv_circ_fwd_cb(dev(inp), int(in_parity), dev(w), int(transposed), dev(bias), int(act),
                                             dev(dact_from), int(dact_kind), dev(out), _C.lattice(shape),
                                             int(ksize), int(Ci), int(Co), B, stream()), "conv_circ_fwd_cb")
        return out
    with _C.timed(f"conv_circ_fwd[{Ci}->{Co}]"):
        check(lib().nfk_conv_circ_fwd(dev(inp), dev(w), int(transposed), dev(bias), dev(in_mask, torch.uint8),
                                      int(in_keep), int(act), dev(dact_from), int(dact_kind), dev(out),
                                      _C.lattice(shape), int(ksize), int(Ci), int(Co), B, stream()),
              "conv_circ_fwd")
    return out


def _conv_dgrad_tc(gpre, w, h, dact_kind, shape, ksize, g_parity=None):
    """Data gradient of a layer on the tensor cores (nfk_convnd_dgrad): d loss / d pre-activation of the layer below
    from `gpre` = d loss / d this layer's output, `w` (Co, Ci, 3, ..) and the post-activation `h` of the layer below
    (tanh) or None.  Returns None where the kernel does not apply (the caller then runs the CUDA-core convolution
    with transposed weights); NFK_DGRAD_TC=0 switches it off, =1 also takes 2-D layers with 8 inputs."""
    Co, Ci = int(w.shape[0]), int(w.shape[1])
    mode = os.environ.get('NFK_DGRAD_TC')
    if (mode == '0' or int(ksize) != 3 or not 2 <= len(shape) <= 4
            or Ci not in (8, 16, 32, 64) or Co > 64 or dact_kind not in (0, _C.ACT['tanh'])):
        return None
    if len(shape) == 2 and Ci == 8 and mode != '1':
        # 2-D, [8, 8] conditioner: the dense 8 -> 8 gradient is break-even with the shared-memory CUDA-core kernel at
        # 64 x 64 and five launches instead of one on small lattices; the checkerboard-sparse gradient of the last layer
        # is faster here (half the MMAs: 1.7 vs 2.05 ms at 64 x 64, B = 4096) once the batch is large enough to fill it
        if g_parity is None or gpre.shape[0] * int(np.prod(shape)) < (1 << 22):
            return None
    lat = _C.lattice(shape)
    B = gpre.shape[0]
    per_sample = int(lib().nfk_convnd_dgrad_workspace(lat, Co, Ci, 1))
    if per_sample <= 0:
        return None
    cap = int(float(os.environ.get("NFK_ND_WORKSPACE_GB", "24")) * 2 ** 30)
    chunk = max(1, min(B, cap // max(per_sample, 1)))
    need = int(lib().nfk_convnd_dgrad_workspace(lat, Co, Ci, chunk))
    workspace = torch.empty((need,), dtype=torch.uint8, device=gpre.device)
    gin = torch.empty((B, Ci) + tuple(shape), dtype=torch.float32, device=gpre.device)
    if dact_kind == 0:
        h = None
    with _C.timed(f"convnd_dgrad[{Co}->{Ci}]"):
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            check(lib().nfk_convnd_dgrad(dev(gpre[lo:hi]), -1 if g_parity is None else int(g_parity), dev(w),
                                         None if h is None else dev(h[lo:hi]), dev(gin[lo:hi]),
                                         Co, Ci, lat, hi - lo, dev(workspace, torch.uint8), need, stream()),
                  "convnd_dgrad")
    return gin


def _conv_layer_bwd_tc(h_in, gpre, w, dact_kind, want_bias, shape, ksize, g_parity=None):
    """Weight, bias and data gradient of one layer with 8 input channels in one call (nfk_convnd_layer_bwd: the gradient
    is reduced and packed into records once for both tensor-core kernels).  Returns (gw, gb, gin) or None where the pair
    does not apply or is not the default (2-D lattices, see _conv_dgrad_tc / _conv_weight_grad)."""
    Co, Ci = int(w.shape[0]), int(w.shape[1])
    if (Ci != 8 or Co > 32 or int(ksize) != 3 or not 3 <= len(shape) <= 4 or shape[-1] % 16 != 0
            or dact_kind not in (0, _C.ACT['tanh'])
            or os.environ.get('NFK_DGRAD_TC') == '0' or os.environ.get('NFK_WGRAD_ND_TC') == '0'
            or os.environ.get('NFK_LAYER_BWD') == '0'):
        return None
    lat = _C.lattice(shape)
    B = gpre.shape[0]
    per_sample = int(lib().nfk_convnd_layer_bwd_workspace(lat, Co, Ci, 1))
    if per_sample <= 0:
        return None
    cap = int(float(os.environ.get("NFK_ND_WORKSPACE_GB", "24")) * 2 ** 30)
    chunk = max(1, min(B, cap // max(per_sample, 1)))
    need = int(lib().nfk_convnd_layer_bwd_workspace(lat, Co, Ci, chunk))
    workspace = torch.empty((need,), dtype=torch.uint8, device=gpre.device)
    gin = torch.empty((B, Ci) + tuple(shape), dtype=torch.float32, device=gpre.device)
    gw = torch.zeros(tuple(w.shape), dtype=torch.float32, device=gpre.device)
    gb = torch.zeros((Co,), dtype=torch.float32, device=gpre.device) if want_bias else None
    with _C.timed(f"convnd_layer_bwd[{Ci}->{Co}]"):
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            check(lib().nfk_convnd_layer_bwd(dev(h_in[lo:hi]), dev(gpre[lo:hi]), -1 if g_parity is None else int(g_parity),
                                             dev(w), int(dact_kind != 0), dev(gin[lo:hi]),
                                             dev(gw), dev(gb), Co, Ci, lat, hi - lo, dev(workspace, torch.uint8), need,
                                             stream()), "convnd_layer_bwd")
    return gw, gb, gin


def _conv_weight_grad(inp, in_mask, in_keep, gpre, w_shape, want_bias, shape, ksize, g_parity=None):
    Co, Ci = w_shape[0], w_shape[1]
    gw = torch.zeros(w_shape, dtype=torch.float32, device=inp.device)
    gb = torch.zeros((Co,), dtype=torch.float32, device=inp.device) if want_bias else None
    # The tensor-core kernel pays for laying the operands out (im2col in shared memory) whatever Co is; it wins
    # where the CUDA-core kernel is slowest: checkerboard-sparse gradients of a wide last layer (8 -> 3K-2).
    # NFK_WGRAD_TC=1 uses it wherever it applies, =0 never.
    mode = os.environ.get('NFK_WGRAD_TC')
    want_tc = mode == '1' or (mode != '0' and g_parity is not None and Co >= 16)
    if in_mask is None and len(shape) == 2 and int(ksize) == 3 and Ci == 8 and Co <= 32 and want_tc:
        # tensor-core kernel (tf32-pair operands, fp32 accumulation); declines geometries it does not cover
        with _C.timed(f"conv2d_wgrad_tc[{Ci}->{Co}]"):
            rc = lib().nfk_conv2d_wgrad_tc(dev(inp), dev(gpre), -1 if g_parity is None else int(g_parity), dev(gw),
                                           dev(gb), int(shape[0]), int(shape[1]), int(Ci), int(Co),
                                           inp.shape[0], stream())
        if rc != _C.EUNSUPPORTED:
            check(rc, "conv2d_wgrad_tc")
            return gw, gb
    nd_mode = os.environ.get('NFK_WGRAD_ND_TC')
    if (in_mask is None and Ci == 8 and Co <= 32 and int(ksize) == 3 and 2 <= len(shape) <= 4 and shape[-1] % 16 == 0
            and (nd_mode == '1' or (nd_mode != '0' and len(shape) >= 3))):
        # sites as the K dimension of tensor-core GEMMs on the site-major records (nfk_convnd_wgrad)
        lat = _C.lattice(shape)
        B = inp.shape[0]
        per_sample = int(lib().nfk_convnd_wgrad_workspace(lat, Co, Ci, 1))
        if per_sample > 0:
            cap = int(float(os.environ.get("NFK_ND_WORKSPACE_GB", "24")) * 2 ** 30)
            chunk = max(1, min(B, cap // max(per_sample, 1)))
            need = int(lib().nfk_convnd_wgrad_workspace(lat, Co, Ci, chunk))
            workspace = torch.empty((need,), dtype=torch.uint8, device=inp.device)
            with _C.timed(f"convnd_wgrad[{Ci}->{Co}]"):
                for lo in range(0, B, chunk):
                    hi = min(B, lo + chunk)
                    check(lib().nfk_convnd_wgrad(dev(inp[lo:hi]), dev(gpre[lo:hi]), dev(gw), dev(gb), Co, Ci, lat, hi - lo,
                                                 dev(workspace, torch.uint8), need, stream()), "convnd_wgrad")
            return gw, gb
    if in_mask is not None and Ci == 1 and len(shape) == 2 and int(ksize) == 3:
        # first layer: apply Mask.split once (one cheap pass) so that the streamed weight-gradient kernel can be used
        inp, in_mask = mask_select(inp, in_mask, in_keep), None
    if g_parity is not None and in_mask is None:
        # gpre lives on one checkerboard partition only: visit just those sites
        with _C.timed(f"conv_circ_bwd_weight_cb[{Ci}->{Co}]"):
            check(lib().nfk_conv_circ_bwd_weight_cb(dev(inp), dev(gpre), int(g_parity), dev(gw), dev(gb),
                                                    _C.lattice(shape), int(ksize), int(Ci), int(Co),
                                                    inp.shape[0], stream()), "conv_circ_bwd_weight_cb")
        return gw, gb
    with _C.timed(f"conv_circ_bwd_weight[{Ci}->{Co}]"):
        check(lib().nfk_conv_circ_bwd_weight(dev(inp), dev(in_mask, torch.uint8), int(in_keep), dev(gpre),
                                             dev(gw), dev(gb), _C.lattice(shape), int(ksize), int(Ci), int(Co),
                                             inp.shape[0], stream()), "conv_circ_bwd_weight")
    return gw, gb


class _ConvStack(torch.autograd.Function):
    """A whole ConvAct stack (conv -> act)* as ONE autograd node.

    forward keeps each layer's post-activation output; backward walks the layers in
    reverse, each step one weight-gradient kernel plus one data-gradient kernel (the
    same circular conv on transposed/flipped weights with act' fused in its epilogue).
    """

    @staticmethod
    @_native
    def forward(ctx, inp, in_mask, in_keep, acts, ksize, n_layers, *params):
        inp = _f32c(inp, "conditioner input")
        weights, biases = params[:n_layers], params[n_layers:]
        shape = tuple(inp.shape[2:])
        hs = [inp]
        for i in range(n_layers):
            w = _f32c(weights[i], "conv weight")
            b = None if biases[i] is None else _f32c(biases[i], "conv bias")
            Co, Ci = w.shape[0], w.shape[1]
            if hs[-1].shape[1] != Ci:
                raise ValueError(f"conv layer {i}: expected {Ci} input channels, got {hs[-1].shape[1]}")
            hs.append(_conv_call(hs[-1], w, 0, b, in_mask if i == 0 else None, in_keep, acts[i], None, 0,
                                 shape, ksize, Ci, Co))
        ctx.save_for_backward(*hs, *[w for w in weights], *( [in_mask] if in_mask is not None else []))
        ctx.cfg = (in_keep, acts, ksize, n_layers, shape, [b is not None for b in biases], in_mask is not None)
        return hs[-1]

    @staticmethod
    @_native
    def backward(ctx, gout):
        in_keep, acts, ksize, n, shape, has_bias, has_mask = ctx.cfg
        saved = ctx.saved_tensors
        hs, weights = saved[:n + 1], saved[n + 1:2 * n + 1]
        in_mask = saved[2 * n + 1] if has_mask else None
        gws, gbs, gin = _conv_stack_backward(hs[:n], weights, has_bias, acts, ksize, shape, in_mask, in_keep,
                                             _f32c(gout, "gout"), ctx.needs_input_grad[0])
        return (gin, None, None, None, None, None, *gws, *gbs)


def _conv_stack_backward(hs, weights, has_bias, acts, ksize, shape, in_mask, in_keep, gpre, need_input_grad,
                         gpre_parity=None, tc_parity=None):
    """Gradients of a ConvAct stack.  hs[i] = input of layer i (hs[0] the stack's input, hs[i] the
    post-activation output of layer i-1), gpre = d loss / d (output of the last layer, which has
    no activation).  Walks the layers in reverse: one weight-gradient kernel plus one
    data-gradient kernel (the same circular conv on transposed/flipped weights with act' fused in
    its epilogue) per layer.  Returns (weight grads, bias grads, input grad or None).
    gpre_parity: gpre vanishes on the sites with coordinate sum % 2 != gpre_parity (2-D checkerboard kernels and the
    tensor-core data gradient skip the known zeros); tc_parity: the same, for the tensor-core data gradient only."""
    n = len(weights)
    if acts[n - 1] != 0:
        raise NotImplementedError("ConvAct with an activation on the output layer: use the unfused path")
    gws, gbs = [None] * n, [None] * n
    for i in reversed(range(n)):
        w = weights[i]
        Co, Ci = w.shape[0], w.shape[1]
        first = (i == 0)
        if not first:
            sp = (tc_parity if tc_parity is not None else gpre_parity) if i == n - 1 else None
            both = _conv_layer_bwd_tc(hs[i], gpre, w.contiguous(), acts[i - 1], has_bias[i], shape, ksize, g_parity=sp)
            if both is not None:
                gws[i], gbs[i], gpre = both
                continue
        gws[i], gbs[i] = _conv_weight_grad(hs[i], in_mask if first else None, in_keep, gpre,
                                           tuple(w.shape), has_bias[i], shape, ksize,
                                           g_parity=gpre_parity if i == n - 1 and n > 1 else None)
        if first and not need_input_grad:
            gpre = None
            break
        # d/d(input of layer i): conv of gpre with w^T (taps flipped); multiply by act'(h_{i})
        if not first:
            sp = (tc_parity if tc_parity is not None else gpre_parity) if i == n - 1 else None
            g_tc = _conv_dgrad_tc(gpre, w.contiguous(), hs[i], acts[i - 1], shape, ksize, g_parity=sp)
            if g_tc is not None:
                gpre = g_tc
                continue
        gpre = _conv_call(gpre, w.contiguous(), 1, None, None, 0, 0, hs[i] if not first else None,
                          acts[i - 1] if not first else 0, shape, ksize, Co, Ci,
                          in_parity=gpre_parity if i == n - 1 else None)
    gin = None
    if gpre is not None:
        gin = gpre if in_mask is None else mask_select(gpre, in_mask, in_keep)
    return gws, gbs, gin


@_native
def conv_stack(inp, weights, biases, acts, ksize, in_mask=None, in_keep=0):
    """ConvAct forward: (B,Ci,*L) -> (B,Co,*L), circular 'same' convs with fused
    activations; `in_mask`/`in_keep` fuse Mask.split into the first layer."""
    acts = tuple(_C.ACT[a] for a in acts)
    if in_mask is not None:
        in_mask = _mask_u8(in_mask)
    return _ConvStack.apply(inp, in_mask, int(in_keep), acts, int(ksize), len(weights), *weights, *biases)


FUSED2D_KNOTS = (4, 5, 6, 8, 10, 12, 16)      # CUDA-core kernel (nfk_fused.cu); needs L1 % 4 == 0
FUSED2D_TC_KNOTS = (4, 5, 6, 8, 10)          # tensor-core kernel (nfk_fused_tc.cu); needs even L0, L1


def fused2d_supported(L0, L1, n_knots=None):
    """Whether nfk_fused2d_step covers a (L0, L1) lattice with an affine (n_knots None) or
    RQ-spline conditioner: mirrors the dispatch in nfk_fused.cu / nfk_fused_tc.cu."""
    tc = (2 <= L1 <= 160 and L0 >= 2 and L0 % 2 == 0 and L1 % 2 == 0          # strip + planes fit one CTA's share
          and (n_knots is None or n_knots in FUSED2D_TC_KNOTS))
    cc = L0 >= 1 and 4 <= L1 <= 512 and L1 % 4 == 0 and (n_knots is None or n_knots in FUSED2D_KNOTS)
    return bool(tc or cc)


@_native
def fused2d_step(x, weights, biases, kind, prm, mask_parity, parity, log0=0, inverse=False):
    """A whole atomic coupling step (ConvAct(1->8->8->P) conditioner + affine / RQ-spline
    transform) in one kernel, forward evaluation only.  x: (B, L0, L1); returns (y, log)."""
    x = _f32c(x, "x")
    if x.dim() != 3:
        raise ValueError("fused2d_step needs a (B, L0, L1) field")
    _no_grad_needed(x, *weights, *[b for b in biases if b is not None])
    log_in = as_log(log0, x)
    B, L0, L1 = x.shape
    w = [_f32c(t, "conv weight") for t in weights]
    b = [None if t is None else _f32c(t, "conv bias") for t in biases]
    y = torch.empty_like(x)
    log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
    if prm is None:
        prm = _C.RqsParams(2, 0.0, 1.0, 0.0, 1.0, 0, 0)
    with _C.timed("fused2d_step"):
        check(lib().nfk_fused2d_step(dev(x), dev(w[0]), dev(b[0]), dev(w[1]), dev(b[1]), dev(w[2]), dev(b[2]),
                                     int(w[0].shape[0]), int(kind), prm, int(mask_parity), int(parity),
                                     int(bool(inverse)), dev(log_in), dev(y), dev(log_out), L0, L1, B, stream()),
              "fused2d_step")
    return y, log_out


def fusednd_supported(shape, n_knots, hidden=8):
    """True when nfk_fusednd_step covers a lattice of this shape (2-D .. 4-D, even extents), knot count
    (None: affine) and hidden width (8, 16, 32 or 64)."""
    shape = tuple(int(v) for v in shape)
    if not 2 <= len(shape) <= 4 or any(v < 2 or v % 2 for v in shape):
        return False
    kind, K = (0, 2) if n_knots is None else (1, int(n_knots))
    return int(lib().nfk_fusednd_workspace(_C.lattice(shape), int(hidden), kind, K, 1)) > 0


@_native
def fusednd_step(x, weights, biases, kind, prm, mask_parity, parity, log0=0, inverse=False):
    """A whole atomic coupling step on a 2-D .. 4-D lattice (ConvAct(1->H->H->P) conditioner with its layers 2
    and 3 on the tensor cores + affine / RQ-spline transform fused into the last layer), forward evaluation
    only.  x: (B, *L); returns (y, log)."""
    x = _f32c(x, "x")
    _no_grad_needed(x, *weights, *[b for b in biases if b is not None])
    log_in = as_log(log0, x)
    B, shape = x.shape[0], tuple(x.shape[1:])
    w = [_f32c(t, "conv weight") for t in weights]
    b = [None if t is None else _f32c(t, "conv bias") for t in biases]
    if prm is None:
        prm = _C.RqsParams(2, 0.0, 1.0, 0.0, 1.0, 0, 0)
    lat = _C.lattice(shape)
    H = int(w[0].shape[0])
    per_sample = int(lib().nfk_fusednd_workspace(lat, H, int(kind), prm.n_knots, 1))
    if per_sample <= 0:
        check(per_sample if per_sample < 0 else _C.EUNSUPPORTED, "fusednd_workspace")
    y = torch.empty_like(x)
    log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
    # the two hidden layers live in the workspace as padded record arrays (about 8 H bytes per site and layer): walk the
    # batch in slices that keep it under NFK_ND_WORKSPACE_GB (default 24 GB) -- samples are independent
    cap = int(float(os.environ.get("NFK_ND_WORKSPACE_GB", "24")) * 2 ** 30)
    chunk = max(1, min(B, cap // max(per_sample, 1)))
    need = int(lib().nfk_fusednd_workspace(lat, H, int(kind), prm.n_knots, chunk))
    workspace = torch.empty((need,), dtype=torch.uint8, device=x.device)     # torch allocations are 512-byte aligned
    with _C.timed("fusednd_step"):
        for lo in range(0, B, chunk):
            hi = min(B, lo + chunk)
            check(lib().nfk_fusednd_step(dev(x[lo:hi]), dev(w[0]), dev(b[0]), dev(w[1]), dev(b[1]), dev(w[2]), dev(b[2]),
                                         H, int(kind), prm, lat, int(mask_parity), int(parity), int(bool(inverse)),
                                         None if log_in is None else dev(log_in[lo:hi]), dev(y[lo:hi]),
                                         dev(log_out[lo:hi]), hi - lo, dev(workspace, torch.uint8), need, stream()),
                  "fusednd_step")
    return y, log_out


class _FusedNdStepTrain(torch.autograd.Function):
    """Training forward of an atomic coupling step on a 2-D .. 4-D lattice through the N-D tensor-core kernels
    (nfk_fusednd_step_train keeps the hidden layers and the conditioner output); backward through the transform's
    VJP kernel and the convolution gradient kernels, as _FusedStepTrain does in 2-D."""

    @staticmethod
    @_native
    def forward(ctx, x, log_in, mask, kind, prm, mask_parity, parity, n_bias, *params):
        w, b = params[:3], params[3:]
        B, shape = x.shape[0], tuple(x.shape[1:])
        H, P = w[0].shape[0], w[2].shape[0]
        lat = _C.lattice(shape)
        need = int(lib().nfk_fusednd_workspace(lat, int(H), int(kind), prm.n_knots, B))
        if need <= 0:
            check(need if need < 0 else _C.EUNSUPPORTED, "fusednd_workspace")
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        h1 = torch.empty((B, H) + shape, dtype=torch.float32, device=x.device)
        h2 = torch.empty((B, H) + shape, dtype=torch.float32, device=x.device)
        out = torch.empty((B, P) + shape, dtype=torch.float32, device=x.device)
        workspace = torch.empty((need,), dtype=torch.uint8, device=x.device)
        with _C.timed("fusednd_step_train"):
            check(lib().nfk_fusednd_step_train(dev(x), dev(w[0]), dev(b[0]), dev(w[1]), dev(b[1]), dev(w[2]), dev(b[2]),
                                               int(H), kind, prm, lat, mask_parity, parity, dev(log_in), dev(y),
                                               dev(log_out), dev(h1), dev(h2), dev(out), B,
                                               dev(workspace, torch.uint8), need, stream()), "fusednd_step_train")
        ctx.save_for_backward(x, h1, h2, out, mask, *w)
        ctx.cfg = (kind, prm, parity, mask_parity, log_in is not None, [t is not None for t in b])
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        x, h1, h2, out, mask, *w = ctx.saved_tensors
        kind, prm, parity, mask_parity, has_log, has_bias = ctx.cfg
        B, shape = x.shape[0], tuple(x.shape[1:])
        V = x[0].numel()
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx, gout = torch.empty_like(x), torch.empty_like(out)
        if kind == 1:
            check(lib().nfk_rqs_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, _C.FROZEN_COPY, prm, dev(gy),
                                    dev(glog), dev(gx), dev(gout), B, V, stream()), "rqs_bwd")
        else:
            check(lib().nfk_affine_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, _C.FROZEN_COPY, dev(gy),
                                       dev(glog), dev(gx), dev(gout), B, V, stream()), "affine_bwd")
        frozen_keep = 0 if parity == 0 else 1          # the conditioner saw the frozen partition only
        acts = (_C.ACT['tanh'], _C.ACT['tanh'], _C.ACT[None])
        # gout is non-zero on the active partition only: sites with (sum of coordinates) % 2 == g_parity
        # (mask bit = (1 - mask_parity + sum) % 2, active <=> bit == (parity == 0))
        g_parity = ((1 if parity == 0 else 0) - 1 + mask_parity) % 2
        gws, gbs, gin = _conv_stack_backward([x.unsqueeze(1), h1, h2], w, has_bias, acts, 3, shape, mask,
                                             frozen_keep, gout, ctx.needs_input_grad[0],
                                             gpre_parity=g_parity if len(shape) == 2 else None, tc_parity=g_parity)
        if gin is not None:
            gx = gx + gin.reshape(x.shape)
        return (gx, glog if has_log else None, None, None, None, None, None, None, *gws, *gbs)


@_native
def fusednd_step_train(x, weights, biases, kind, prm, mask, mask_parity, parity, log0=0):
    """Differentiable atomic coupling step on a 2-D .. 4-D lattice: tensor-core forward, kernel-by-kernel backward."""
    x = _f32c(x, "x")
    log_in = as_log(log0, x)
    w = [_f32c(t, "conv weight") for t in weights]
    b = [None if t is None else _f32c(t, "conv bias") for t in biases]
    if prm is None:
        prm = _C.RqsParams(2, 0.0, 1.0, 0.0, 1.0, 0, 0)
    return _FusedNdStepTrain.apply(x, log_in, _mask_u8(mask), int(kind), prm, int(mask_parity), int(parity),
                                   sum(t is not None for t in b), *w, *b)


class _FusedStepTrain(torch.autograd.Function):
    """Forward of a training step through the tensor-core fused kernel (which also stores the
    hidden layers and the conditioner output); backward through the transform's VJP kernel and
    the convolution gradient kernels -- the conditioner is never evaluated layer by layer."""

    @staticmethod
    @_native
    def forward(ctx, x, log_in, mask, kind, prm, mask_parity, parity, n_bias, *params):
        w, b = params[:3], params[3:]
        B, L0, L1 = x.shape
        P = w[2].shape[0]
        y = torch.empty_like(x)
        log_out = torch.empty((B,), dtype=torch.float32, device=x.device)
        h1 = torch.empty((B, 8, L0, L1), dtype=torch.float32, device=x.device)
        h2 = torch.empty((B, 8, L0, L1), dtype=torch.float32, device=x.device)
        out = torch.empty((B, P, L0, L1), dtype=torch.float32, device=x.device)
        with _C.timed("fused2d_step_train"):
            check(lib().nfk_fused2d_step_train(dev(x), dev(w[0]), dev(b[0]), dev(w[1]), dev(b[1]), dev(w[2]), dev(b[2]),
                                               8, kind, prm, mask_parity, parity, dev(log_in), dev(y), dev(log_out),
                                               dev(h1), dev(h2), dev(out), L0, L1, B, stream()), "fused2d_step_train")
        ctx.save_for_backward(x, h1, h2, out, mask, *w)
        ctx.cfg = (kind, prm, parity, mask_parity, log_in is not None, [t is not None for t in b])
        return y, log_out

    @staticmethod
    @_native
    def backward(ctx, gy, glog):
        x, h1, h2, out, mask, *w = ctx.saved_tensors
        kind, prm, parity, mask_parity, has_log, has_bias = ctx.cfg
        B, L0, L1 = x.shape
        V = L0 * L1
        gy, glog = _f32c(gy, "gy"), _f32c(glog, "glog")
        gx, gout = torch.empty_like(x), torch.empty_like(out)
        if kind == 1:
            with _C.timed("rqs_bwd"):
                check(lib().nfk_rqs_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, _C.FROZEN_COPY, prm, dev(gy),
                                        dev(glog), dev(gx), dev(gout), B, V, stream()), "rqs_bwd")
        else:
            check(lib().nfk_affine_bwd(dev(x), dev(out), dev(mask, torch.uint8), parity, _C.FROZEN_COPY, dev(gy),
                                       dev(glog), dev(gx), dev(gout), B, V, stream()), "affine_bwd")
        frozen_keep = 0 if parity == 0 else 1          # the conditioner saw the frozen partition only
        acts = (_C.ACT['tanh'], _C.ACT['tanh'], _C.ACT[None])
        # gout is non-zero on the active partition only: sites with (row + col) % 2 == g_parity
        # (mask bit = (1 - mask_parity + row + col) % 2, active <=> bit == (parity == 0))
        active_val = 1 if parity == 0 else 0
        g_parity = (active_val - 1 + mask_parity) % 2
        gws, gbs, gin = _conv_stack_backward([x.unsqueeze(1), h1, h2], w, has_bias, acts, 3, (L0, L1), mask,
                                             frozen_keep, gout, ctx.needs_input_grad[0], gpre_parity=g_parity)
        if gin is not None:
            gx = gx + gin.reshape(x.shape)
        return (gx, glog if has_log else None, None, None, None, None, None, None, *gws, *gbs)


def fused2d_train_supported(L0, L1, n_knots=None):
    """Geometry of the tensor-core kernel (the only one with a training forward)."""
    return (2 <= L1 <= 160 and L0 >= 2 and L0 % 2 == 0 and L1 % 2 == 0
            and (n_knots is None or n_knots in FUSED2D_TC_KNOTS))


@_native
def fused2d_step_train(x, weights, biases, kind, prm, mask, mask_parity, parity, log0=0):
    """Differentiable atomic coupling step: fused forward, kernel-by-kernel backward."""
    x = _f32c(x, "x")
    log_in = as_log(log0, x)
    w = [_f32c(t, "conv weight") for t in weights]
    b = [None if t is None else _f32c(t, "conv bias") for t in biases]
    if prm is None:
        prm = _C.RqsParams(2, 0.0, 1.0, 0.0, 1.0, 0, 0)
    return _FusedStepTrain.apply(x, log_in, _mask_u8(mask), int(kind), prm, int(mask_parity), int(parity),
                                 sum(t is not None for t in b), *w, *b)


# ---------------------------------------------------------------------------- mcmc
@_native
def metropolis_scan(logq, logp, log_u, ref_state):
    """Sequential accept/reject on the device (mcmc.py:304-328).  `ref_state` is a
    float64[2] CUDA tensor {ref, has_ref}, updated in place.  Returns
    (accept uint8[B], idx int64[B], n_accept int64[1])."""
    logq, logp = _f32c(logq, "logq"), _f32c(logp, "logp")
    B = logq.shape[0]
    accept = torch.empty((B,), dtype=torch.uint8, device=logq.device)
    idx = torch.empty((B,), dtype=torch.int64, device=logq.device)
    n_acc = torch.empty((1,), dtype=torch.int64, device=logq.device)
    check(lib().nfk_metropolis_scan(dev(logq), dev(logp), dev(log_u, torch.float64), dev(ref_state, torch.float64),
                                    dev(accept, torch.uint8), dev(idx, torch.int64), dev(n_acc, torch.int64),
                                    B, stream()), "metropolis_scan")
    return accept, idx, n_acc


@_native
def metropolis_rates(logqp, perm, log_u):
    """Acceptance rates of R shuffled chains (mcmc.py:117-124): logqp float64[N], perm int64[R, N] (or
    None), log_u float64[R, N] -> float64[R], all on the device; one launch, no synchronisation."""
    R, N = log_u.shape
    if logqp.numel() != N or (perm is not None and tuple(perm.shape) != (R, N)):
        raise ValueError("metropolis_rates: logqp [N], perm [R, N], log_u [R, N]")
    rates = torch.empty((R,), dtype=torch.float64, device=logqp.device)
    check(lib().nfk_metropolis_rates(dev(logqp, torch.float64), dev(perm, torch.int64), dev(log_u, torch.float64),
                                     dev(rates, torch.float64), N, R, stream()), "metropolis_rates")
    return rates


@_native
def gather_rows(src, idx, prev=None):
    """dst[i] = src[idx[i]] (idx >= 0) or prev (idx < 0)  (mcmc.py:67-75)."""
    src = _f32c(src, "src")
    B = src.shape[0]
    row = src.numel() // max(B, 1)
    dst = torch.empty_like(src)
    if prev is not None:
        prev = _f32c(prev, "prev")
        if prev.numel() != row:
            raise ValueError("gather_rows: `prev` must be one row")
    check(lib().nfk_gather_rows(dev(src), dev(idx, torch.int64), dev(prev), dev(dst), B, row, stream()),
          "gather_rows")
    return dst
