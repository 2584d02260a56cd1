"""GPU parity of the PSD block: PSDBlock_ / MeanFieldNet_ / FFTNet_ through
the package API (cuFFT + the nfk_psd_* / nfk_sample_* kernels of the C ABI) against the
reference's golden outputs, the numpy oracle, and size-independent properties at full size.
Tolerance as everywhere: 1e-5 * max(|ref|, 1)."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import nf_oracle as O
from test_oracle_psd import CASES, psd_params

import normflow__b200 as nf  # noqa: F401  (default device / dtype)
from normflow__b200 import Model, _ops
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import (ModuleList_, ConvAct, AffineCoupling_, DistConvertor_, Identity_,
                               FFTNet_, MeanFieldNet_, PSDBlock_, IPSD)
from normflow__b200.prior import NormalPrior
from test_gpu_parity import cu, close, close_grad, DEV, ACTION

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ kernels on their own
@pytest.mark.parametrize("shape", [(8, 4), (6,), (4, 6, 3), (64, 33), (16, 16, 16, 9)])
@pytest.mark.parametrize("inverse", [False, True])
def test_psd_weights_kernel_vs_oracle(shape, inverse):
    rng = np.random.RandomState(11)
    ipsd = (rng.rand(*shape) * 3 + 0.05).astype(np.float32)
    t = cu(ipsd).requires_grad_(True)
    w, logj = _ops.psd_weights(t, inverse=inverse)
    wref = 1 / ipsd.astype(np.float64) ** 0.5
    lref = O.fftnet_log_jacobian(wref)
    if inverse:
        wref, lref = 1 / wref, -lref
    close(w, wref)
    close(logj, lref)
    assert logj.dim() == 0
    # adjoint: L = sum(w r) + c logj
    r = rng.randn(*shape).astype(np.float32)
    (g,) = torch.autograd.grad((w * cu(r)).sum() + 0.7 * logj, t)
    mult = np.full(shape, 2.0)
    mult[..., 0] -= 1
    mult[..., -1] -= 1
    sgn = 0.5 if inverse else -0.5
    gref = sgn * (r * wref + 0.7 * mult) / ipsd.astype(np.float64)
    close_grad(g, gref)


@pytest.mark.parametrize("B,shape", [(1, (4, 3)), (5, (8, 4)), (7, (6,)), (300, (16, 9)), (3, (4, 6, 3)),
                                     (2050, (64, 33))])
@pytest.mark.parametrize("replace", [False, True])
def test_psd_scale_kernel_and_adjoint(B, shape, replace):
    rng = np.random.RandomState(12)
    X = (rng.randn(B, *shape) + 1j * rng.randn(B, *shape)).astype(np.complex64)
    w = (rng.rand(*shape) + 0.5).astype(np.float32)
    z = rng.randn(B).astype(np.float32)
    Xt = torch.as_tensor(X).to(DEV).requires_grad_(True)
    wt = cu(w).requires_grad_(True)
    zt = cu(z).requires_grad_(True) if replace else None
    Y = _ops.psd_scale(Xt, wt, zt, 3.0)
    ref = X.astype(np.complex128) * w
    if replace:
        ref.reshape(B, -1)[:, 0] = 3.0 * z
    got = Y.detach().cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1)
    G = (rng.randn(B, *shape) + 1j * rng.randn(B, *shape)).astype(np.complex64)
    Gt = torch.as_tensor(G).to(DEV)
    L = (Y.real * Gt.real + Y.imag * Gt.imag).sum()
    grads = torch.autograd.grad(L, [Xt, wt] + ([zt] if replace else []))
    gx_ref = G.astype(np.complex128) * w
    gw_ref = (X.real.astype(np.float64) * G.real + X.imag.astype(np.float64) * G.imag).sum(0)
    if replace:
        gx_ref.reshape(B, -1)[:, 0] = 0
        gw_ref.reshape(-1)[0] = 0
        close_grad(grads[2], 3.0 * G.reshape(B, -1)[:, 0].real)
    gx = grads[0].cpu().numpy()
    assert np.abs(gx - gx_ref).max() <= 1e-6 * max(np.abs(gx_ref).max(), 1)
    close_grad(grads[1], gw_ref, tol=2e-6 * np.sqrt(B))
    # in-place evaluation path (no gradient recorded) gives the same numbers
    with torch.no_grad():
        X2 = torch.as_tensor(X).to(DEV)
        Y2 = _ops.psd_scale(X2, wt.detach(), None if zt is None else zt.detach(), 3.0)
    assert Y2.data_ptr() == X2.data_ptr()
    assert torch.equal(Y2, Y.detach())


def test_psd_scale_rejects_bad_input():
    X = torch.zeros(2, 4, 3, dtype=torch.complex64, device=DEV)
    with pytest.raises(ValueError):
        _ops.psd_scale(X, torch.ones(4, 4, device=DEV))
    with pytest.raises(TypeError):
        _ops.psd_scale(X.to(torch.complex128), torch.ones(4, 3, device=DEV))
    with pytest.raises(RuntimeError):
        _ops.psd_scale(X.cpu(), torch.ones(4, 3, device=DEV))
    with pytest.raises(ValueError):
        _ops.psd_scale(X, torch.ones(4, 3, device=DEV), torch.zeros(3, device=DEV))


@pytest.mark.parametrize("B,shape", [(1, (1,)), (9, (5,)), (33, (8, 8)), (5, (64, 64)), (3, (16, 16, 16)),
                                     (4, (7, 3))])
def test_sample_mean_and_shift(B, shape):
    rng = np.random.RandomState(13)
    x = rng.randn(B, *shape).astype(np.float32)
    d = rng.randn(B).astype(np.float32)
    xt, dt = cu(x).requires_grad_(True), cu(d).requires_grad_(True)
    m = _ops.sample_mean(xt)
    close(m, x.astype(np.float64).reshape(B, -1).mean(1), tol=2e-6)
    y = _ops.sample_shift(xt, dt)
    close(y, x.astype(np.float64) + d.reshape(-1, *[1] * len(shape)), tol=1e-6)
    r = rng.randn(B, *shape).astype(np.float32)
    c = rng.randn(B).astype(np.float32)
    gx, gd = torch.autograd.grad((y * cu(r)).sum() + (m * cu(c)).sum(), [xt, dt])
    V = int(np.prod(shape))
    close_grad(gx, r + (c / V).reshape(-1, *[1] * len(shape)))
    close_grad(gd, r.astype(np.float64).reshape(B, -1).sum(1), tol=2e-6)


# ------------------------------------------------------------------ knot table (SplineNet.make_spline)
def _knots_torch64(wx, wy, wd, xlim, ylim):
    """modules.py:369-391 with torch float64 ops on the CPU (differentiable)."""
    def coord(w, lim):
        p = torch.softmax(w, 0)
        left = torch.cat([p.new_zeros(1), torch.cumsum(p, 0)])
        return lim[0] + (lim[1] - lim[0]) * left, (lim[1] - lim[0]) * p
    kx, bx = coord(wx, xlim)
    ky, by = coord(wy, ylim)
    if wd is None:
        s = by / bx
        kd = torch.cat([s[:1], 0.5 * (s[1:] + s[:-1]), s[-1:]])
    else:
        kd = torch.nn.functional.softplus(wd, beta=float(np.log(2)))
    return torch.stack([kx, ky, kd, xlim[1] - kx, ylim[1] - ky])


@pytest.mark.parametrize("K,smooth,xlim,ylim", [(2, True, (0, 1), (0, 1)), (2, False, (0.5, 1), (0.5, 1)),
                                                (10, False, (0, 1), (0, 1)), (10, True, (0.5, 1), (0.5, 1)),
                                                (50, True, (-2, 3), (0, 7)), (200, False, (0, 1), (-1, 1))])
def test_knot_table_kernel_and_adjoint(K, smooth, xlim, ylim):
    g = torch.Generator('cpu').manual_seed(K)
    wx = torch.randn(K - 1, generator=g, device='cpu')
    wy = torch.randn(K - 1, generator=g, device='cpu')
    wd = None if smooth else torch.randn(K, generator=g, device='cpu') * 3
    if wd is not None:
        wd[0] = 40.0        # beyond the softplus threshold
    r = torch.randn(5, K, generator=g, device='cpu')
    ref_in = [t.double().requires_grad_(True) for t in (wx, wy) + (() if smooth else (wd,))]
    ref = _knots_torch64(ref_in[0], ref_in[1], None if smooth else ref_in[2], xlim, ylim)
    ref_g = torch.autograd.grad((ref * r.double()).sum(), ref_in)
    dev_in = [t.to(DEV).requires_grad_(True) for t in (wx, wy) + (() if smooth else (wd,))]
    table = _ops.knot_table(dev_in[0], dev_in[1], None if smooth else dev_in[2], xlim, ylim)
    close(table, ref.detach().numpy(), tol=2e-7 if K < 100 else 1e-6)
    assert float(table[0, -1]) == float(np.float32(xlim[1])) and float(table[3, -1]) == 0.0
    got_g = torch.autograd.grad((table * r.to(DEV)).sum(), dev_in)
    for a, b in zip(got_g, ref_g):
        close_grad(a, b.numpy(), tol=2e-6)


# ------------------------------------------------------------------ modules against the reference's outputs
def _build_case(g, tag):
    lat = tuple(int(v) for v in g[f"{tag}_lat_shape"])
    has_mf, ignore, alone = CASES[tag]
    _, _, ipsd_w, _ = psd_params(g, tag)
    K = len(ipsd_w[0]) + 1
    kw = dict(smooth=True) if ipsd_w[2] is None else {}
    ipsd_net = IPSD(K, logy=torch.zeros(2), ignore_zeromode=ignore, **kw)
    ff = FFTNet_(lat, ipsd_net, ignore_zeromode=ignore)
    if alone:
        net = ff
    else:
        mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True) if has_mf \
            else Identity_()
        net = PSDBlock_(mfnet_=mf, fftnet_=ff)
    names = [str(n) for n in g[f"{tag}_param_names"]]
    assert [n for n, _ in net.named_parameters()] == names       # state_dict compatible
    sd = {n: cu(g[f"{tag}_w_{n}"]) for n in names}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("lat_k2") for k in missing)
    return net, ff


@pytest.mark.parametrize("tag", sorted(CASES))
def test_psd_modules_golden(tag):
    g = load_golden("psd")
    net, ff = _build_case(g, tag)
    close(ff.norm_lat_k2, g[f"{tag}_norm_lat_k2"], tol=2e-7)
    close(ff.max_lat_k2, g[f"{tag}_max_lat_k2"], tol=2e-7)
    close(ff.ipsd, g[f"{tag}_ipsd"])
    x = cu(g[f"{tag}_x"]).requires_grad_(True)
    B = x.shape[0]
    y, logJ = net(x)
    close(y, g[f"{tag}_y"])
    logJ_full = logJ if logJ.dim() > 0 else logJ * torch.ones(B, device=DEV)
    close(logJ_full, g[f"{tag}_logJ"])
    # reference-autograd gradients of L = sum(y r) + sum(logJ c)
    L = (y * cu(g[f"{tag}_r"])).sum() + (logJ_full * cu(g[f"{tag}_c"])).sum()
    params = dict(net.named_parameters())
    grads = torch.autograd.grad(L, [x] + list(params.values()))
    close_grad(grads[0], g[f"{tag}_gx"])
    for (n, _), gr in zip(params.items(), grads[1:]):
        close_grad(gr, g[f"{tag}_grad_{n}"], tol=2e-5)
    with torch.no_grad():
        xb, lb = net.backward(cu(g[f"{tag}_y"]), log0=cu(g[f"{tag}_logJ"]))
        close(xb, g[f"{tag}_rt_x"], tol=2e-5)
        close(lb, g[f"{tag}_rt_log"], tol=2e-5)
        yi, li = net.backward(cu(g[f"{tag}_x"]))
        close(yi, g[f"{tag}_inv_y"], tol=2e-5)
        close(li if li.dim() > 0 else li * torch.ones(B, device=DEV), g[f"{tag}_inv_logJ"], tol=2e-5)


def test_psd_hack_parts_golden():
    g = load_golden("psd")
    net, _ = _build_case(g, "ex2d")
    with torch.no_grad():
        (xm, l0), (ymf, lmf), (yfft, lfft), (y, ltot) = net._hack(cu(g["ex2d_x"]))
    assert l0 == 0
    close(xm, g["ex2d_x_mean"], tol=1e-6)
    close(ymf, g["ex2d_y_mf"])
    close(lmf, g["ex2d_logJ_mf"])
    close(yfft, g["ex2d_y_fft"])
    close(lfft, g["ex2d_logJ_fft"])
    close(y, g["ex2d_y"])
    close(ltot, g["ex2d_logJ"])


def test_log_jacobian_method_matches_kernel():
    ff = FFTNet_.build((8, 6), knots_len=7)
    with torch.no_grad():
        for p in ff.parameters():
            p.add_(0.2 * torch.randn_like(p))
        w, logj = ff.weights()
        close(ff.log_jacobian(w), float(logj))
        w_inv, logj_inv = ff.weights(inverse=True)
        close(w * w_inv, np.ones((8, 4)), tol=1e-6)
        close(logj_inv, -float(logj))


def test_meanfieldnet_on_a_whole_field_vs_oracle():
    """rvol=None path (meanfield_.py:26-32): x + (dc(mean * rvol) / rvol - mean)."""
    g = load_golden("psd")
    net, _ = _build_case(g, "ex2d")
    mf = net.mfnet_
    mf_w, fs, _, _ = psd_params(g, "ex2d")
    x = g["ex2d_x"]
    rvol = np.prod(x.shape[1:]) ** 0.5
    mean = x.mean(axis=(1, 2)).reshape(-1, 1, 1)
    for inverse in (False, True):
        ref_m, ref_l = O.meanfieldnet(mean, 0, rvol, mf_w, symmetric=True, final_scale=fs, inverse=inverse)
        with torch.no_grad():
            y, l = (mf.backward if inverse else mf.forward)(cu(x))
        close(y, x + (ref_m - mean), tol=2e-5 if inverse else 1e-5)
        close(l, ref_l, tol=2e-5 if inverse else 1e-5)
    # gradient flows through mean and shift kernels
    xt = cu(x).requires_grad_(True)
    y, l = mf(xt)
    (gx,) = torch.autograd.grad(y.sum() + l.sum(), xt)
    assert torch.isfinite(gx).all()
    stack = mf._hack(cu(x))
    close(stack[0][0], mean.ravel(), tol=1e-6)


def test_generic_composition_equals_spectral_path():
    """PSDBlock_ with a fftnet_ that is not an FFTNet_ instance runs the literal composition
    (mean kernel, nets, shift kernel); wrapping the same FFTNet_ must give the same numbers."""
    g = load_golden("psd")
    net, ff = _build_case(g, "ex2d")

    class Wrapped(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, log0=0):
            return self.inner.forward(x, log0)

        def backward(self, x, log0=0):
            return self.inner.backward(x, log0)

    generic = PSDBlock_(mfnet_=net.mfnet_, fftnet_=Wrapped(ff))
    x = cu(g["ex2d_x"])
    with torch.no_grad():
        for fn_a, fn_b in ((net.forward, generic.forward), (net.backward, generic.backward)):
            ya, la = fn_a(x)
            yb, lb = fn_b(x)
            close(yb, ya.cpu().numpy(), tol=3e-6)
            close(lb, la.cpu().numpy(), tol=3e-6)
    close(generic(x)[0], g["ex2d_y"])


def test_errors_match_intent():
    with pytest.raises(ValueError):
        FFTNet_.build((4, 5))                      # odd last axis
    ff = FFTNet_.build((4, 4))
    with pytest.raises(ValueError):
        ff(torch.zeros(2, 4, 6, device=DEV))       # wrong lattice
    with pytest.raises(RuntimeError):
        ff(torch.zeros(2, 4, 4, device="cpu"))     # no CPU path


def test_fftnet_without_batch_axis_and_density():
    ff = FFTNet_.build((4, 6), knots_len=4)
    x = torch.randn(3, 4, 6, device=DEV)
    with torch.no_grad():
        y, l = ff(x)
        y0, l0 = ff(x[0])
        assert y0.shape == (4, 6) and l0.dim() == 0
        close(y0, y[0].cpu().numpy(), tol=1e-6)
        nf.nn.Module_.propagate_density = True
        try:
            d = ff.create_density(l)
            assert d.shape == (4, 6)
            close(d.sum(), float(l))
        finally:
            nf.nn.Module_.propagate_density = False


def test_transfer_to_finer_lattice():
    """FFTNet_.transfer / IPSD.apply_scale (fftflow_.py:187-212, 253-258): logy shifted by
    log(1/scale_factor) * (ndim, ndim-2), spline weights copied, buffers rebuilt for the new shape."""
    ff = FFTNet_.build((8, 8), knots_len=6, ignore_zeromode=True)
    with torch.no_grad():
        ff.ipsd_net.weights_x.add_(0.3 * torch.randn(5, device=DEV))
    new = ff.transfer(scale_factor=2, shape=(16, 16))
    assert new.lat_shape == (16, 16) and new.norm_lat_k2.shape == (16, 9) and new.ignore_zeromode
    a = np.log(0.5)
    close(new.ipsd_net.logy, ff.ipsd_net.logy.detach().cpu().numpy() + np.array([2 * a, 0.0]), tol=1e-6)
    assert torch.equal(new.ipsd_net.weights_x, ff.ipsd_net.weights_x)
    assert new.ipsd_net.weights_x.data_ptr() != ff.ipsd_net.weights_x.data_ptr()
    blk = PSDBlock_(mfnet_=MeanFieldNet_.build(knots_len=4, symmetric=True), fftnet_=ff)
    blk2 = blk.transfer(scale_factor=2, shape=(16, 16))
    with torch.no_grad():
        y, l = blk2(torch.randn(2, 16, 16, device=DEV))
    assert y.shape == (2, 16, 16) and l.shape == (2,)
    close(ff.infrared_mass, np.exp(0.5 * float(ff.ipsd_net.logy[0].detach())), tol=1e-6)


# ------------------------------------------------------------------ the whole example net
def _example_net(lat, knots=12, seed=0):
    torch.manual_seed(seed)
    mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
    ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
    conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, padding_mode='circular',
                conv_dim=len(lat), acts=('tanh', 'tanh', None), bias=False)
    return ModuleList_([
        PSDBlock_(mfnet_=mf, fftnet_=ff),
        DistConvertor_(knots, symmetric=True, smooth=True),
        AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
        DistConvertor_(knots, symmetric=True, smooth=True)])


def test_model_psd_affine_golden():
    """examples/scalar_affine.py's net: block outputs, logq / logp / loss, every parameter's gradient
    and the inverse flow against the reference run recorded in model_psd_affine.npz."""
    g = load_golden("model_psd_affine")
    lat = tuple(int(v) for v in g["lat_shape"])
    net_ = _example_net(lat)
    names = [str(n) for n in g["param_names"]]
    assert [n for n, _ in net_.named_parameters()] == names
    missing, unexpected = net_.load_state_dict({n: cu(g["w_" + n]) for n in names}, strict=False)
    assert not unexpected and all(("lat_k2" in k) or ("mask" in k) for k in missing)
    prior, action = NormalPrior(shape=lat), ScalarPhi4Action(**ACTION)
    model = Model(net_=net_, prior=prior, action=action)
    x = cu(g["x"])
    stack = net_.hack(x, log0=0)
    for i, (yi, li) in enumerate(stack[1:]):
        close(yi, g[f"blk{i}_y"])
        close(li, g[f"blk{i}_logJ"])
    y, logJ = net_(x)
    logq, logp = prior.log_prob(x) - logJ, -action(y)
    close(logq, g["logq"])
    close(logp, g["logp"])
    loss = model.fit.calc_kl_mean(logq, logp)
    close(loss, g["loss"])
    grads = torch.autograd.grad(loss, list(net_.parameters()))
    for n, gr in zip(names, grads):
        close_grad(gr, g["g_" + n], tol=2e-5)
    with torch.no_grad():
        xb, lb = net_.backward(cu(g["y"]), log0=cu(g["blk3_logJ"]))
    close(xb, g["inv_x"], tol=3e-5)
    # the round trip cancels a log-Jacobian of magnitude |logJ|: residual measured against that scale
    scale = max(float(np.abs(g["blk3_logJ"]).max()), 1.0)
    assert float(np.abs(lb.double().cpu().numpy() - g["inv_log"]).max()) <= 2e-6 * scale


@pytest.mark.parametrize("lat,B", [((64, 64), 4096), ((32, 32, 32), 64), ((16, 16, 16, 16), 16), ((128, 128), 257)])
def test_psd_block_full_size_properties(lat, B):
    """Size-independent checks at the BASELINE lattice sizes: forward o backward = identity and the
    log-Jacobians cancel; the fluctuation part is linear; the lattice average of the output is the
    mean-field net's output; log J of the spectral part equals the oracle's closed form."""
    torch.manual_seed(5)
    blk = PSDBlock_(mfnet_=MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True),
                    fftnet_=FFTNet_.build(lat, knots_len=10, ignore_zeromode=True))
    with torch.no_grad():
        for n, p in blk.named_parameters():
            p.add_((0.05 if n.endswith('logy') else 0.3) * torch.randn_like(p))
        x = torch.randn(B, *lat, device=DEV)
        y, logJ = blk(x)
        xb, lb = blk.backward(y, logJ)
        scale = float(x.abs().max())
        assert float((xb - x).abs().max()) <= 2e-5 * scale
        assert float(lb.abs().max()) <= 1e-5 * max(float(logJ.abs().max()), 1)
        # output mean == mean-field output
        V = int(np.prod(lat))
        x_mean = _ops.sample_mean(x)
        y_mf, logJ_mf = blk.mfnet_(x_mean.reshape(-1, *[1] * len(lat)), rvol=V ** 0.5)
        close(_ops.sample_mean(y), y_mf.reshape(-1).cpu().numpy(), tol=1e-5)
        # spectral log J against the oracle's closed form on the kernel's own ipsd
        ipsd = blk.fftnet_.ipsd.double().cpu().numpy()
        close((logJ - logJ_mf), O.fftnet_log_jacobian(1 / ipsd ** 0.5) * np.ones(B))
        # linearity of the fluctuation part: F(a x1 + b x2) = a F(x1) + b F(x2)
        ff = blk.fftnet_
        x2 = torch.randn(B, *lat, device=DEV)
        lhs = ff(0.5 * x - 1.5 * x2)[0]
        rhs = 0.5 * ff(x)[0] - 1.5 * ff(x2)[0]
        assert float((lhs - rhs).abs().max()) <= 1e-5 * float(rhs.abs().max())
        # against the oracle on a few samples
        ref, _ = O.fftnet(x[:2].double().cpu().numpy(), 0, ipsd)
        close(ff(x[:2])[0], ref)


def test_example_net_trains_eager_and_graph():
    """Model.fit on the example net (PSDBlock_ first) lowers the loss; the CUDA-graph trainer follows
    the eager one on the same seeds."""
    lat = (8, 8)
    losses = {}
    for mode in (False, True):
        torch.manual_seed(21)
        np.random.seed(21)
        net_ = _example_net(lat, knots=8, seed=3)
        model = Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(**ACTION))
        model.device_handler.to(DEV)
        model.fit.cuda_graph = mode
        model.fit(n_epochs=60, batch_size=256, checkpoint_dict=dict(print_stride=30, print_batch_size=128),
                  hyperparam=dict(lr=0.005, weight_decay=0.0))
        hist = np.array(model.fit.train_history['loss'])
        assert len(hist) == 60 and np.isfinite(hist).all()
        assert hist[-5:].mean() < hist[:5].mean() - 1.0
        losses[mode] = hist
    np.testing.assert_allclose(losses[True][:5], losses[False][:5], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(losses[True], losses[False], rtol=5e-2, atol=0.5)


def test_example_scripts_run(tmp_path, capsys):
    """examples/scalar_zerodim.py and examples/scalar_affine.py (this package's versions of the reference's two
    example scripts) train, write a snapshot in the reference's format and pass the backward sanity check."""
    import importlib.util
    import os
    from conftest import ROOT

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    torch.manual_seed(1)
    np.random.seed(1)
    model = load("scalar_zerodim").main(n_epochs=300, batch_size=1024, print_stride=100)
    assert np.mean(model.fit.train_history['loss'][-20:]) < -1.05
    snap = tmp_path / "affine.E0.tar"
    model = load("scalar_affine").main(lat_shape=(8, 8), n_epochs=40, batch_size=128, print_stride=20,
                                       snapshot_path=str(snap), save_every=20, knots2_len=12, knots4_len=12)
    hist = np.array(model.fit.train_history['loss'])
    assert len(hist) == 40 and np.isfinite(hist).all() and hist[-5:].mean() < hist[:5].mean()
    assert (tmp_path / "affine.E20.tar").exists() and (tmp_path / "affine.E40.tar").exists()
    out = capsys.readouterr().out
    assert "Sanity check is OK" in out
