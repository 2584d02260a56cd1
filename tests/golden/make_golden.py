"""Generate golden fixtures from the REFERENCE itself (jkomijani/normflow_).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is pure Python; it is imported from /root/reference/src through a
temporary symlink named `normflow_ref`, with the `np.product` shim numpy>=2 needs
(header table).  It runs on CPU in its default float64.  Inputs are
drawn in float32 and widened, so the CUDA tests can feed bit-identical values.
The outputs land in tests/golden/*.npz (committed).  Nothing at test/bench time
reads /root/reference.
"""

import os
import sys
import tempfile
import warnings

import numpy as np

warnings.filterwarnings("ignore")
np.product = np.prod  # removed in numpy 2; the reference still calls it

HERE = os.path.dirname(os.path.abspath(__file__))
_tmp = tempfile.mkdtemp()
os.symlink("/root/reference/src", os.path.join(_tmp, "normflow_ref"))
sys.path.insert(0, _tmp)

import torch  # noqa: E402
import normflow_ref as nf  # noqa: E402  (sets default dtype float64)
from normflow_ref.mask import EvenOddMask, AlongAxesEvenOddMask  # noqa: E402
from normflow_ref.action import ScalarPhi4Action  # noqa: E402
from normflow_ref.prior import NormalPrior  # noqa: E402
from normflow_ref.nn import (ModuleList_, ConvAct, AffineCoupling_, ShiftCoupling_,  # noqa: E402
                             RQSplineCoupling_, DistConvertor_, Identity_,
                             FFTNet_, MeanFieldNet_, PSDBlock_, MultiRQSplineCoupling_,
                             CntrAffineCoupling_, CntrRQSplineCoupling_, CntrShiftCoupling_)
from normflow_ref.lib.spline import RQSpline  # noqa: E402
from normflow_ref.mcmc.mcmc import Metropolis, MCMCSampler, BlockedMCMCSampler  # noqa: E402
from normflow_ref import Model  # noqa: E402

assert torch.get_default_dtype() == torch.float64

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)


def f32(t):
    """Round to float32-representable values, keep float64 storage."""
    return t.float().double()


def randn32(*shape, seed):
    g = torch.Generator('cpu').manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).double()


def npy(t):
    if not torch.is_tensor(t):   # e.g. ShiftCoupling_ hands log0=0 through untouched
        return np.asarray(t, dtype=np.float64)
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def round_params(module):
    for p in module.parameters():
        p.data = f32(p.data)


# -----------------------------------------------------------------------------
def gen_masks():
    out = {}
    cases = [((4, 4), 0, None), ((4, 4), 1, None), ((1,), 0, None), ((5,), 1, None),
             ((3, 5), 0, None), ((4, 6), 0, 1), ((4, 4, 4), 1, None), ((2, 3, 4), 0, 0),
             ((3, 2, 2, 3), 0, None), ((4, 4, 4, 4), 1, 3)]
    for i, (shape, parity, mu) in enumerate(cases):
        m = EvenOddMask(shape=shape, parity=parity, exclude_mu=mu)
        out[f"eo_{i}_mask"] = npy(m._mask)
        out[f"eo_{i}_cmask"] = npy(m._c_mask)
        out[f"eo_{i}_meta"] = np.array([parity, -1 if mu is None else mu] + list(shape))
    for i, (shape, parity, mu) in enumerate([((4, 6), 0, 0), ((4, 6), 1, 1), ((3, 4, 5), 0, 2)]):
        m = AlongAxesEvenOddMask(shape=shape, parity=parity, mu=mu)
        out[f"aa_{i}_mask"] = npy(m._mask)
        out[f"aa_{i}_meta"] = np.array([parity, mu] + list(shape))
    save("masks", **out)


def gen_action():
    out = {}
    for i, shape in enumerate([(1,), (2,), (7,), (8, 8), (5, 6), (4, 4, 4), (3, 4, 5), (4, 4, 4, 4), (2, 3, 2, 3)]):
        cfgs = randn32(6, *shape, seed=100 + i).requires_grad_(True)
        act = ScalarPhi4Action(**ACTION)
        S = act(cfgs)
        w = randn32(6, seed=200 + i)
        (g,) = torch.autograd.grad((S * w).sum(), cfgs)
        out[f"a{i}_cfgs"] = npy(cfgs)
        out[f"a{i}_S"] = npy(S)
        out[f"a{i}_density"] = npy(act.action_density(cfgs))
        out[f"a{i}_gS"] = npy(w)
        out[f"a{i}_gcfgs"] = npy(g)
    # zero-dim config-1 action (kappa=0)
    cfgs = randn32(128, 1, seed=31)
    out["zd_cfgs"] = npy(cfgs)
    out["zd_S"] = npy(ScalarPhi4Action(kappa=0, m_sq=-1.2, lambd=0.5)(cfgs))
    # a != 1
    cfgs = randn32(3, 4, 4, seed=32)
    out["lat_cfgs"] = npy(cfgs)
    out["lat_S"] = npy(ScalarPhi4Action(kappa=0.5, m_sq=0.3, lambd=0.2, a=0.5)(cfgs))
    save("action", **out)


def gen_prior():
    out = {}
    x = randn32(5, 4, 6, seed=41)
    prior = NormalPrior(shape=(4, 6))
    out["std_x"], out["std_logr"] = npy(x), npy(prior.log_prob(x))
    loc = f32(randn32(4, 6, seed=42) * 0.3)
    scale = f32(torch.rand(4, 6, generator=torch.Generator('cpu').manual_seed(43)) + 0.5)
    prior = NormalPrior(loc=loc, scale=scale)
    out["gen_x"], out["gen_loc"], out["gen_scale"] = npy(x), npy(loc), npy(scale)
    out["gen_logr"] = npy(prior.log_prob(x))
    save("prior", **out)


def gen_spline():
    """RQSpline with explicit knots along axis 1, all extrapolation modes."""
    out = {}
    B, K, V = 3, 6, 40
    g = torch.Generator('cpu').manual_seed(51)
    wx = torch.rand(B, K - 1, V, generator=g) + 0.2
    wy = torch.rand(B, K - 1, V, generator=g) + 0.2
    kx = f32(torch.cat([torch.zeros(B, 1, V), torch.cumsum(wx / wx.sum(1, keepdim=True), 1)], 1) * 2 - 1)
    ky = f32(torch.cat([torch.zeros(B, 1, V), torch.cumsum(wy / wy.sum(1, keepdim=True), 1)], 1) * 3 - 1.5)
    kd = f32(torch.rand(B, K, V, generator=g) * 1.5 + 0.25)
    x = f32((torch.rand(B, 1, V, generator=g) * 4 - 2))   # covers outside [-1, 1]
    # put a few points exactly on knots (searchsorted right=False edge)
    x[0, 0, :K] = kx[0, :, 0]
    kx[0, :, :K] = kx[0, :, :1]
    out.update(kx=npy(kx), ky=npy(ky), kd=npy(kd), x=npy(x))
    modes = {"none": {}, "lin": dict(left='linear', right='linear'),
             "linleft": dict(left='linear'), "linright": dict(right='linear'),
             "antileft": dict(left='anti'), "antiright": dict(right='anti'),
             "antilin": dict(left='anti', right='linear')}
    for name, extrap in modes.items():
        sp = RQSpline(knots_x=kx, knots_y=ky, knots_d=kd, knots_axis=1, extrap=extrap)
        y, gr = sp(x, grad=True)
        out[f"{name}_y"], out[f"{name}_g"] = npy(y), npy(gr)
        out[f"{name}_nknots"] = np.array(sp.knots_x.shape[1])
        # inverse evaluated on the forward image (in-range by construction)
        xi, gi = sp.backward(y, grad=True)
        out[f"{name}_xinv"], out[f"{name}_ginv"] = npy(xi), npy(gi)
    # smooth derivatives, 1-D shared knots, anti-left (the DistConvertor_ smooth case)
    k1x, k1y = kx[1, :, 3].contiguous(), ky[1, :, 3].contiguous()
    sp = RQSpline(knots_x=k1x, knots_y=k1y, knots_d=None, extrap=dict(left='anti'))
    xs = f32(torch.rand(200, generator=g) * 5 - 3.5)
    ys, gs = sp(xs, grad=True)
    out.update(s1_kx=npy(k1x), s1_ky=npy(k1y), s1_x=npy(xs), s1_y=npy(ys), s1_g=npy(gs),
               s1_kd=npy(sp.knots_d), s1_kxaug=npy(sp.knots_x))
    save("spline", **out)


def gen_spline_extra():
    """(a) `periodic` extrapolation (spline.py:502-508, 518-524) of a shared-knot RQSpline whose end
    derivatives are zero: values and derivatives in range, in the mirrored range and beyond it.
    (b) RQSplineCoupling_ with FIXED knots_x and / or knots_y (couplings_.py:246-258): outputs, log J,
    inverse and reference-autograd gradients."""
    out = {}
    g = torch.Generator('cpu').manual_seed(77)
    K = 7
    wx, wy = torch.rand(K - 1, generator=g) + 0.3, torch.rand(K - 1, generator=g) + 0.3
    kx = f32(torch.cat([torch.zeros(1), torch.cumsum(wx / wx.sum(), 0)]) * 3 - 1)         # [-1, 2]
    ky = f32(torch.cat([torch.zeros(1), torch.cumsum(wy / wy.sum(), 0)]) * 2 + 0.5)       # [0.5, 2.5]
    kd = f32(torch.rand(K, generator=g) * 1.2 + 0.3)
    kd[0], kd[-1] = 0.0, 0.0
    x = f32(torch.rand(400, generator=g) * 11 - 5)      # [-5, 6]: both mirror images ([-4,-1], [2,5]) and beyond
    out.update(per_kx=npy(kx), per_ky=npy(ky), per_kd=npy(kd), per_x=npy(x))
    for name, extrap in {"perleft": dict(left='periodic'), "perright": dict(right='periodic'),
                         "perboth": dict(left='periodic', right='periodic'),
                         "perleft_antiright": dict(left='periodic', right='anti')}.items():
        sp = RQSpline(knots_x=kx, knots_y=ky, knots_d=kd, extrap=extrap)
        y, gr = sp(x, grad=True)
        out[f"{name}_y"], out[f"{name}_g"] = npy(y), npy(gr)
    # fixed knots in the coupling
    shape, B, m = (6, 4), 3, 6
    mask = EvenOddMask(shape=shape)
    fx = f32(torch.tensor([-3.0, -1.7, -0.4, 0.3, 1.9, 3.0]))
    fy = f32(torch.tensor([-2.5, -1.0, -0.2, 0.9, 1.4, 2.5]))
    out.update(fix_kx=npy(fx), fix_ky=npy(fy), fix_shape=np.array(shape))
    cases = {"fixx": (dict(knots_x=fx), 2 * m - 1, dict(left='linear', right='linear')),
             "fixy": (dict(knots_y=fy), 2 * m - 1, {}),
             "fixxy": (dict(knots_x=fx, knots_y=fy), m, {})}
    for ci, (tag, (kw, P, extrap)) in enumerate(cases.items()):
        torch.manual_seed(300 + ci)
        nets = [ConvAct(1, P, 3, conv_dim=2, hidden_sizes=[4], acts=['tanh', None], bias=True) for _ in range(2)]
        for n in nets:
            round_params(n)
        cpl = RQSplineCoupling_(nets, mask=mask, xlim=(-3, 3), ylim=(-2.5, 2.5), extrap=extrap, **kw)
        xx = f32(randn32(B, *shape, seed=400 + ci) * 1.6).requires_grad_(True)
        y, logJ = cpl(xx, log0=0)
        loss = (y ** 2).sum(dim=(1, 2)).mean() - logJ.mean()
        params = list(cpl.parameters())
        grads = torch.autograd.grad(loss, [xx] + params)
        out.update({f"{tag}_x": npy(xx), f"{tag}_y": npy(y), f"{tag}_logJ": npy(logJ), f"{tag}_loss": npy(loss),
                    f"{tag}_gx": npy(grads[0])})
        gi = 1
        for k, net in enumerate(cpl.nets):
            for key, val in conv_layers(net).items():
                out[f"{tag}_step{k}_{key}"] = val
            for pname, p_ in net.named_parameters():
                out[f"{tag}_step{k}_grad_{pname}"] = npy(grads[gi])
                gi += 1
        with torch.no_grad():
            xb, lb = cpl.backward(y.detach(), log0=logJ.detach())
        out[f"{tag}_inv_x"], out[f"{tag}_inv_log"] = npy(xb), npy(lb)
    save("spline_extra", **out)


def conv_layers(net):
    """Export ConvAct weights in standard (Co, Ci, *k) shape."""
    ws = {}
    i = 0
    for m in net:
        if hasattr(m, "weight") and not isinstance(m, (torch.nn.Tanh,)):
            ws[f"w{i}"] = npy(m.weight)
            if getattr(m, "bias", None) is not None:
                ws[f"b{i}"] = npy(m.bias)
            if hasattr(m, "_conv_lower_dim"):
                ws[f"wlower{i}"] = npy(m._conv_lower_dim.weight)
            i += 1
    return ws


def gen_coupling(name, shape, B, blocks, seed, hidden=(8, 8), bias=False, acts=None, per_step=True):
    """blocks: list of ('affine'|'shift'|'rqs', n_steps).  Records every block's
    output (ModuleList_.hack), the inverse, and reference-autograd gradients of
    loss = mean(logr - logJ + S) w.r.t. x and every parameter."""
    torch.manual_seed(seed)
    nd = len(shape)
    K = 10
    if acts is None:
        acts = (*['tanh'] * len(hidden), None)
    mask = EvenOddMask(shape=shape)
    nets_, meta = [], []
    for kind, n_steps in blocks:
        P = {'affine': 2, 'shift': 1, 'rqs': 3 * K - 2}[kind]
        nets = [ConvAct(1, P, 3, conv_dim=nd, hidden_sizes=list(hidden), acts=acts, bias=bias)
                for _ in range(n_steps)]
        for n in nets:
            round_params(n)
        if kind == 'affine':
            nets_.append(AffineCoupling_(nets, mask=mask))
        elif kind == 'shift':
            nets_.append(ShiftCoupling_(nets, mask=mask))
        else:
            nets_.append(RQSplineCoupling_(nets, mask=mask, xlim=(-5, 5), ylim=(-5, 5),
                                           extrap=dict(left='linear', right='linear')))
        meta.append((kind, n_steps))
    net_ = ModuleList_(nets_)
    x = randn32(B, *shape, seed=seed + 1)
    if any(k == 'rqs' for k, _ in blocks):
        x = x * 2.0   # push a few sites outside xlim so linear extrapolation is hit
        x = f32(x)
    x.requires_grad_(True)
    prior = NormalPrior(shape=shape)
    action = ScalarPhi4Action(**ACTION)
    out = dict(x=npy(x), shape=np.array(shape), mask=npy(mask._mask))
    stack = net_.hack(x, log0=0)
    for i, (xi, li) in enumerate(stack[1:]):
        out[f"blk{i}_y"], out[f"blk{i}_logJ"] = npy(xi), npy(li)
    y, logJ = stack[-1]
    if not torch.is_tensor(logJ):
        logJ = torch.zeros(B) + logJ
    logr = prior.log_prob(x)
    S = action(y)
    loss = (logr - logJ + S).mean()
    params = list(net_.parameters())
    grads = torch.autograd.grad(loss, [x] + params)
    out.update(y=npy(y), logJ=npy(logJ), logr=npy(logr), S=npy(S), loss=npy(loss), gx=npy(grads[0]))
    # weights + grads, addressed as blk{i}_step{k}_{w,b}{layer}
    gi = 1
    for bi, cpl in enumerate(net_):
        for k, net in enumerate(cpl.nets):
            for key, val in conv_layers(net).items():
                out[f"blk{bi}_step{k}_{key}"] = val
            for pname, p in net.named_parameters():
                out[f"blk{bi}_step{k}_grad_{pname}"] = npy(grads[gi])
                gi += 1
    assert gi == len(grads)
    # per-step conditioner outputs for the first block (kernel-only parity)
    with torch.no_grad():
        cpl = net_[0]
        parts = list(cpl.mask.split(x))
        for k, net in enumerate(cpl.nets if per_step else []):   # (per_step=False: large lattices, keep the file small)
            p = k % 2
            o = net(parts[1 - p].unsqueeze(1))
            out[f"blk0_step{k}_out"] = npy(o)
            out[f"blk0_step{k}_xactive"] = npy(parts[p])
            parts[p], _ = cpl.atomic_forward(x_active=parts[p], x_frozen=parts[1 - p],
                                             parity=p, net=net, log0=0)
            out[f"blk0_step{k}_fx"] = npy(parts[p])
        # inverse of the whole flow
        xb, lb = net_.backward(y.detach(), log0=logJ.detach())
        out["inv_x"], out["inv_log"] = npy(xb), npy(lb)
    out["blocks"] = np.array([f"{k}:{n}" for k, n in meta])
    out["hidden"] = np.array(hidden)
    out["acts"] = np.array(['none' if a is None else a for a in acts])
    save(name, **out)


def gen_rqs_kernel_only():
    """Kernel-only parity for one RQS atomic step: explicit conditioner output
    `out = 0.5 randn`, forward values + reference-autograd gradients
    w.r.t. x_active and out for random upstream weights; plus the inverse."""
    res = {}
    for tag, shape, B, extrap in [("lin", (8, 8), 3, dict(left='linear', right='linear')),
                                  ("none", (6,), 4, {}),
                                  ("mixed", (4, 4, 4), 2, dict(left='linear'))]:
        K = 10
        mask = EvenOddMask(shape=shape)
        cpl = RQSplineCoupling_([torch.nn.Identity()], mask=mask, xlim=(-5, 5), ylim=(-4, 6),
                                extrap=extrap)
        x = f32(randn32(B, *shape, seed=61) * 2.5)
        out_raw = f32(randn32(B, 3 * K - 2, *shape, seed=4321) * 0.5)
        for parity in (0, 1):
            xa = cpl.mask.purify(x, parity).clone().requires_grad_(True)
            o = out_raw.clone().requires_grad_(True)
            sp = cpl.make_spline(o)
            fx, g = sp(xa.unsqueeze(1), grad=True)
            fx, g = fx.squeeze(1), g.squeeze(1)
            fx = cpl.mask.purify(fx, parity)
            logJ = cpl.sum_density(cpl.mask.purify(torch.log(g), parity))
            r = randn32(B, *shape, seed=62)
            c = randn32(B, seed=63)
            L = (fx * r).sum() + (logJ * c).sum()
            gxa, go = torch.autograd.grad(L, [xa, o])
            xi, gi = sp.backward(fx.detach().unsqueeze(1), grad=True)
            res.update({f"{tag}_p{parity}_fx": npy(fx), f"{tag}_p{parity}_logJ": npy(logJ),
                        f"{tag}_p{parity}_gx": npy(gxa), f"{tag}_p{parity}_gout": npy(go),
                        f"{tag}_p{parity}_xinv": npy(cpl.mask.purify(xi.squeeze(1), parity)),
                        f"{tag}_p{parity}_loginv": npy(cpl.sum_density(
                            cpl.mask.purify(torch.log(gi.squeeze(1)), parity)))})
        res.update({f"{tag}_x": npy(x), f"{tag}_out": npy(out_raw), f"{tag}_r": npy(r),
                    f"{tag}_c": npy(c), f"{tag}_mask": npy(mask._mask)})
    save("rqs_kernel", **res)


def gen_affine_kernel_only():
    res = {}
    shape, B = (8, 8), 3
    mask = EvenOddMask(shape=shape)
    cpl = AffineCoupling_([torch.nn.Identity()], mask=mask)
    x = randn32(B, *shape, seed=71)
    out_raw = f32(randn32(B, 2, *shape, seed=72) * 0.7)
    r, c = randn32(B, *shape, seed=73), randn32(B, seed=74)
    for parity in (0, 1):
        xa = cpl.mask.purify(x, parity).clone().requires_grad_(True)
        o = out_raw.clone().requires_grad_(True)
        fx, logJ = cpl.atomic_forward(x_active=xa, x_frozen=x, parity=parity,
                                      net=lambda _: o, log0=0)
        L = (fx * r).sum() + (logJ * c).sum()
        gxa, go = torch.autograd.grad(L, [xa, o])
        xi, li = cpl.atomic_backward(x_active=fx.detach(), x_frozen=x, parity=parity,
                                     net=lambda _: o.detach(), log0=0)
        res.update({f"p{parity}_fx": npy(fx), f"p{parity}_logJ": npy(logJ), f"p{parity}_gx": npy(gxa),
                    f"p{parity}_gout": npy(go), f"p{parity}_xinv": npy(xi), f"p{parity}_loginv": npy(li)})
    res.update(x=npy(x), out=npy(out_raw), r=npy(r), c=npy(c), mask=npy(mask._mask))
    save("affine_kernel", **res)


def gen_distconv():
    res = {}
    for tag, kw, shape, B in [("zd_sym", dict(symmetric=True), (1,), 128),
                              ("lat_sym_smooth", dict(symmetric=True, smooth=True), (4, 4), 6),
                              ("lat_asym", dict(symmetric=False), (6,), 9)]:
        torch.manual_seed(81)
        K = 10
        net_ = DistConvertor_(K, **kw)
        for p in net_.parameters():
            p.data = f32(torch.randn_like(p) * 0.5)
        x = randn32(B, *shape, seed=82)
        if tag == "lat_asym":
            x = f32(x * 3.0)   # reach the tails: |x| up to ~9
        x.requires_grad_(True)
        y, logJ = net_(x)
        r, c = randn32(B, *shape, seed=83), randn32(B, seed=84)
        L = (y * r).sum() + (logJ * c).sum()
        params = list(net_.parameters())
        grads = torch.autograd.grad(L, [x] + params)
        with torch.no_grad():
            xb, lb = net_.backward(y.detach(), log0=logJ.detach())
        sp = net_.spline_layer_
        res.update({f"{tag}_x": npy(x), f"{tag}_y": npy(y), f"{tag}_logJ": npy(logJ),
                    f"{tag}_r": npy(r), f"{tag}_c": npy(c), f"{tag}_gx": npy(grads[0]),
                    f"{tag}_inv_x": npy(xb), f"{tag}_inv_log": npy(lb),
                    f"{tag}_wx": npy(sp.weights_x), f"{tag}_wy": npy(sp.weights_y)})
        names = [n for n, _ in net_.named_parameters()]
        for n, gval in zip(names, grads[1:]):
            res[f"{tag}_grad_{n}"] = npy(gval)
        if sp.weights_d is not None:
            res[f"{tag}_wd"] = npy(sp.weights_d)
        res[f"{tag}_param_names"] = np.array(names)
    save("distconv", **res)


def gen_mcmc():
    res = {}
    rng = np.random.RandomState(5)
    logqp = rng.randn(64).astype(np.float32).astype(np.float64) * 0.8
    np.random.seed(7)
    res["u_first"] = np.random.rand(64)
    np.random.seed(7)
    st = Metropolis.calc_accept_status(logqp)
    res.update(logqp=logqp, status_noref=st, ind_noref=Metropolis.calc_accept_indices(st))
    np.random.seed(8)
    res["u_second"] = np.random.rand(64)
    np.random.seed(8)
    st = Metropolis.calc_accept_status(logqp, logqp_ref=-0.3)
    res.update(status_ref=st, ind_ref=Metropolis.calc_accept_indices(st), ref=np.array(-0.3))

    # two consecutive MCMCSampler._accept_reject_step calls (chain state carried over)
    class _M:  # minimal stand-in for Model: _accept_reject_step never touches it
        pass
    sampler = MCMCSampler(_M())
    for call in range(2):
        y = randn32(16, 4, 4, seed=90 + call)
        logq = randn32(16, seed=92 + call)
        logp = f32(logq + randn32(16, seed=94 + call) * 0.7)
        np.random.seed(100 + call)
        u = np.random.rand(16)
        np.random.seed(100 + call)
        yo, lqo, lpo = sampler._accept_reject_step(y.clone(), logq.clone(), logp.clone())
        res.update({f"c{call}_y": npy(y), f"c{call}_logq": npy(logq), f"c{call}_logp": npy(logp),
                    f"c{call}_u": u, f"c{call}_yo": npy(yo), f"c{call}_logqo": npy(lqo),
                    f"c{call}_logpo": npy(lpo),
                    f"c{call}_accept_rate": np.array(sampler.history.accept_rate[-1])})
    save("mcmc", **res)


def gen_conv():
    """Circular ConvAct stacks alone, every supported dimension (conv parity),
    including Conv4d in its stored (lower-dim) weight layout and a biased one."""
    res = {}
    for tag, shape, hidden, P, bias in [("d1", (9,), (4,), 3, True), ("d2", (6, 5), (8, 8), 28, False),
                                        ("d3", (4, 3, 5), (4,), 2, True), ("d4", (3, 4, 3, 4), (3,), 2, True)]:
        torch.manual_seed(111)
        nd = len(shape)
        net = ConvAct(1, P, 3, conv_dim=nd, hidden_sizes=list(hidden),
                      acts=(*['tanh'] * len(hidden), None), bias=bias)
        round_params(net)
        x = randn32(2, 1, *shape, seed=112).requires_grad_(True)
        o = net(x)
        r = randn32(*o.shape, seed=113)
        grads = torch.autograd.grad((o * r).sum(), [x] + list(net.parameters()))
        res.update({f"{tag}_x": npy(x), f"{tag}_out": npy(o), f"{tag}_r": npy(r), f"{tag}_gx": npy(grads[0])})
        for key, val in conv_layers(net).items():
            res[f"{tag}_{key}"] = val
        for (pname, _), gval in zip(net.named_parameters(), grads[1:]):
            res[f"{tag}_grad_{pname}"] = npy(gval)
        res[f"{tag}_hidden"] = np.array(hidden)
    save("conv", **res)


def gen_model_zero_dim():
    """Config 1 end to end: posterior.sample__ with the prior draw pinned, and
    one Fitter-style loss + gradient."""
    torch.manual_seed(121)
    net_ = DistConvertor_(10, symmetric=True)
    for p in net_.parameters():
        p.data = f32(torch.randn_like(p) * 0.3)
    prior = NormalPrior(shape=1)
    action = ScalarPhi4Action(kappa=0, m_sq=-1.2, lambd=0.5)
    model = Model(net_=net_, prior=prior, action=action)
    x = randn32(128, 1, seed=122)
    logr = prior.log_prob(x)
    y, logJ = net_(x)
    logq = logr - logJ
    logp = -action(y)
    loss = model.fit.calc_kl_mean(logq, logp)
    grads = torch.autograd.grad(loss, list(net_.parameters()))
    res = dict(x=npy(x), y=npy(y), logq=npy(logq), logp=npy(logp), loss=npy(loss))
    for (n, p), gval in zip(net_.named_parameters(), grads):
        res[f"w_{n}"] = npy(p)
        res[f"g_{n}"] = npy(gval)
    save("model_zero_dim", **res)

# -----------------------------------------------------------------------------
def _randomise(net_, scale=0.3):
    """float32-representable random parameters; IPSD's `logy` keeps its built value plus
    a small offset (it sets the overall scale of the spectrum)."""
    for n, p in net_.named_parameters():
        if n.endswith('logy'):
            p.data = f32(p.data + 0.1 * torch.randn_like(p))
        else:
            p.data = f32(torch.randn_like(p) * scale)


def _psd_record(tag, blk, x, res, seed):
    """forward, hack parts, inverse and reference-autograd gradients of one PSDBlock_ / FFTNet_."""
    B = x.shape[0]
    x.requires_grad_(True)
    y, logJ = blk(x)
    logJ_full = logJ if (torch.is_tensor(logJ) and logJ.dim() > 0) else torch.zeros(B) + logJ
    r, c = randn32(*x.shape, seed=seed + 1), randn32(B, seed=seed + 2)
    L = (y * r).sum() + (logJ_full * c).sum()
    names = [n for n, _ in blk.named_parameters()]
    grads = torch.autograd.grad(L, [x] + list(blk.parameters()))
    with torch.no_grad():
        xb, lb = blk.backward(y.detach(), log0=logJ_full.detach())
        # the inverse direction on an independent input (not just the round trip)
        yi, li = blk.backward(x.detach())
    li = li if (torch.is_tensor(li) and li.dim() > 0) else torch.zeros(B) + li
    res.update({f"{tag}_x": npy(x), f"{tag}_y": npy(y), f"{tag}_logJ": npy(logJ_full),
                f"{tag}_r": npy(r), f"{tag}_c": npy(c), f"{tag}_gx": npy(grads[0]),
                f"{tag}_rt_x": npy(xb), f"{tag}_rt_log": npy(lb),
                f"{tag}_inv_y": npy(yi), f"{tag}_inv_logJ": npy(li),
                f"{tag}_param_names": np.array(names)})
    for n, p, g in zip(names, blk.parameters(), grads[1:]):
        res[f"{tag}_w_{n}"] = npy(p)
        res[f"{tag}_grad_{n}"] = npy(g)
    fft = blk.fftnet_ if hasattr(blk, 'fftnet_') else blk
    res[f"{tag}_ipsd"] = npy(fft.ipsd)
    res[f"{tag}_norm_lat_k2"] = npy(fft.norm_lat_k2)
    res[f"{tag}_max_lat_k2"] = npy(fft.max_lat_k2)
    res[f"{tag}_lat_shape"] = np.array(fft.lat_shape)
    if hasattr(blk, '_hack'):
        with torch.no_grad():
            stack = blk._hack(x.detach())
        (xm, _), (ymf, lmf), (yfft, lfft), _ = stack
        res.update({f"{tag}_x_mean": npy(xm), f"{tag}_y_mf": npy(ymf), f"{tag}_logJ_mf": npy(lmf),
                    f"{tag}_y_fft": npy(yfft), f"{tag}_logJ_fft": npy(lfft)})


def gen_psd():
    """PSDBlock_ / MeanFieldNet_ / FFTNet_ (first block of examples/scalar_affine.py:71-77)."""
    res = {}
    # the example's block on a 2-D lattice with unequal sides
    torch.manual_seed(131)
    lat = (8, 6)
    mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
    ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
    blk = PSDBlock_(mfnet_=mf, fftnet_=ff)
    _randomise(blk)
    _psd_record("ex2d", blk, randn32(5, *lat, seed=132), res, seed=133)
    # 3-D, identity mean field (knots0_len <= 1 in the example), zero mode kept, non-unit a
    torch.manual_seed(141)
    lat = (4, 6, 4)
    ff = FFTNet_.build(lat, knots_len=6, eff_mass2=0.7, eff_kappa=1.3, a=0.5)
    blk = PSDBlock_(mfnet_=Identity_(), fftnet_=ff)
    _randomise(blk)
    _psd_record("id3d", blk, randn32(3, *lat, seed=142), res, seed=143)
    # FFTNet_ on its own, 1-D, identity spline (knots_len < 2), zero mode kept
    torch.manual_seed(151)
    lat = (10,)
    ff = FFTNet_.build(lat, knots_len=1, eff_mass2=0.5)
    _randomise(ff)
    _psd_record("fft1d", ff, randn32(4, *lat, seed=152), res, seed=153)
    # FFTNet_ on its own, 2-D, asymmetric mean-field variant not involved; 4x4 with zero mode ignored
    torch.manual_seed(161)
    lat = (4, 4)
    ff = FFTNet_.build(lat, knots_len=5, ignore_zeromode=True)
    _randomise(ff)
    _psd_record("fft2d", ff, randn32(6, *lat, seed=162), res, seed=163)
    save("psd", **res)


def gen_model_psd_affine():
    """The whole net of examples/scalar_affine.py:61-114 on 8x8 (smaller spline sizes): PSDBlock_,
    DistConvertor_, AffineCoupling_ x 4, DistConvertor_; loss and gradients of Fitter.step."""
    torch.manual_seed(171)
    lat = (8, 8)
    mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
    ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
    conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3,
                padding_mode='circular', conv_dim=2, acts=('tanh', 'tanh', None), bias=False)
    net_ = ModuleList_([
        PSDBlock_(mfnet_=mf, fftnet_=ff),
        DistConvertor_(12, symmetric=True, smooth=True),
        AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
        DistConvertor_(12, symmetric=True, smooth=True)])
    for n, p in net_.named_parameters():
        if n.endswith('logy'):
            p.data = f32(p.data + 0.1 * torch.randn_like(p))
        elif p.dim() > 1:
            p.data = f32(p.data)                     # conv weights: torch default init
        else:
            p.data = f32(torch.randn_like(p) * 0.3)
    prior = NormalPrior(shape=lat)
    action = ScalarPhi4Action(**ACTION)
    model = Model(net_=net_, prior=prior, action=action)
    x = randn32(6, *lat, seed=172)
    logr = prior.log_prob(x)
    stack = net_.hack(x, log0=0)
    y, logJ = stack[-1]
    logq, logp = logr - logJ, -action(y)
    loss = model.fit.calc_kl_mean(logq, logp)
    grads = torch.autograd.grad(loss, list(net_.parameters()))
    res = dict(x=npy(x), y=npy(y), logq=npy(logq), logp=npy(logp), loss=npy(loss),
               lat_shape=np.array(lat), param_names=np.array([n for n, _ in net_.named_parameters()]))
    for i, (xi, li) in enumerate(stack[1:]):
        res[f"blk{i}_y"], res[f"blk{i}_logJ"] = npy(xi), npy(li)
    for (n, p), g in zip(net_.named_parameters(), grads):
        res[f"w_{n}"] = npy(p)
        res[f"g_{n}"] = npy(g)
    with torch.no_grad():
        xb, lb = net_.backward(y.detach(), log0=logJ.detach())
    res["inv_x"], res["inv_log"] = npy(xb), npy(lb)
    save("model_psd_affine", **res)


def gen_snapshot():
    """On-disk formats written BY THE REFERENCE: a training snapshot
    {"MODEL_STATE", "EPOCHS_RUN"} saved by Fitter._save_snapshot (_normflowcore.py:236-247) after
    3 CPU epochs of the scalar_affine example net on 8x8, the same weights as a text blob
    (ModuleList_.get_weights_blob, nn/_core.py:108-111), the flow's output on a fixed input, and the
    state_dict layout of a 4-D net (Conv4d keys)."""
    import shutil
    torch.manual_seed(181)
    lat = (8, 8)
    mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
    ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
    conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3,
                padding_mode='circular', conv_dim=2, acts=('tanh', 'tanh', None), bias=False)
    net_ = ModuleList_([
        PSDBlock_(mfnet_=mf, fftnet_=ff),
        DistConvertor_(12, symmetric=True, smooth=True),
        AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
        DistConvertor_(12, symmetric=True, smooth=True)])
    model = Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(**ACTION))
    tmp = tempfile.mkdtemp()
    model.fit(n_epochs=3, batch_size=32, save_every=3,
              checkpoint_dict=dict(print_stride=10, snapshot_path=os.path.join(tmp, "ref_affine.E0.tar")))
    shutil.copy(os.path.join(tmp, "ref_affine.E3.tar"), os.path.join(HERE, "ref_affine.E3.tar"))
    with open(os.path.join(HERE, "ref_affine_blob.txt"), "w") as fh:
        fh.write(net_.get_weights_blob() + "\n")
    x = randn32(4, *lat, seed=182)
    with torch.no_grad():
        y, logJ = net_(x)
    res = dict(x=npy(x), y=npy(y), logJ=npy(logJ),
               keys=np.array(list(net_.state_dict().keys())),
               shapes=np.array([",".join(map(str, v.shape)) for v in net_.state_dict().values()]),
               dtypes=np.array([str(v.dtype) for v in net_.state_dict().values()]))
    # 4-D net: layout only
    torch.manual_seed(183)
    lat4 = (4, 4, 4, 4)
    conv4 = dict(in_channels=1, out_channels=2, hidden_sizes=[4], kernel_size=3, padding_mode='circular',
                 conv_dim=4, acts=('tanh', None), bias=True)
    net4 = ModuleList_([AffineCoupling_([ConvAct(**conv4) for _ in range(2)], mask=EvenOddMask(shape=lat4))])
    res["keys4d"] = np.array(list(net4.state_dict().keys()))
    res["shapes4d"] = np.array([",".join(map(str, v.shape)) for v in net4.state_dict().values()])
    save("snapshot_layout", **res)
    print("wrote ref_affine.E3.tar", os.path.getsize(os.path.join(HERE, "ref_affine.E3.tar")), "bytes")


class SqueezeChannel(torch.nn.Module):
    """Conditioner wrapper for multi-component data: Coupling_.preprocess_fz hands the frozen field
    over as (B, 1, S, *L); the convolution wants the S components as its input channels."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return self.net(x.squeeze(1))


def _record_flow(tag, net_, x, res, seed, inv_in):
    B = x.shape[0]
    x.requires_grad_(True)
    y, logJ = net_(x)
    if not (torch.is_tensor(logJ) and logJ.dim() > 0):
        logJ = torch.zeros(B) + logJ
    r, c = randn32(*x.shape, seed=seed + 1), randn32(B, seed=seed + 2)
    L = (y * r).sum() + (logJ * c).sum()
    names = [n for n, _ in net_.named_parameters()]
    grads = torch.autograd.grad(L, [x] + list(net_.parameters()), allow_unused=True)
    # the inverse direction on in-range points only: the reference's inverse loses all digits on the
    # linear-extrapolation segments (a2 ~ 1e-16 in spline.py:262-281), where there is nothing to pin
    with torch.no_grad():
        xb, lb = net_.backward(inv_in, log0=c)
    res.update({f"{tag}_x": npy(x), f"{tag}_y": npy(y), f"{tag}_logJ": npy(logJ), f"{tag}_r": npy(r),
                f"{tag}_c": npy(c), f"{tag}_gx": npy(grads[0]), f"{tag}_inv_in": npy(inv_in),
                f"{tag}_inv_x": npy(xb), f"{tag}_inv_log": npy(lb), f"{tag}_param_names": np.array(names)})
    for n, p, g in zip(names, net_.parameters(), grads[1:]):
        res[f"{tag}_w_{n}"] = npy(p)
        res[f"{tag}_grad_{n}"] = npy(g if g is not None else torch.zeros_like(p))


def gen_rank4_couplings():
    """MultiRQSplineCoupling_ (couplings_.py:279-436) and the controlled couplings
    (cntr_couplings_.py) on small lattices."""
    res = {}
    lat = (4, 6)
    mask = EvenOddMask(shape=lat)
    K = 5
    lin = dict(left='linear', right='linear')
    for tag, xlims, ylims, extraps in [
            ("multi_uniform", [(-3, 3)] * 2, [(-3, 3)] * 2, [lin, lin]),
            ("multi_mixed", [(-3, 3), (-2, 2.5), (-4, 4)], [(-3, 3), (-2, 2.5), (-4, 4)], [lin, lin, {}])]:
        torch.manual_seed(191)
        S = len(xlims)
        nets = [SqueezeChannel(ConvAct(S, S * (3 * K - 2), 3, hidden_sizes=[4], acts=('tanh', None)))
                for _ in range(3)]
        for n in nets:
            round_params(n)
        cpl = MultiRQSplineCoupling_(nets, mask=mask, xlims=xlims, ylims=ylims, extraps=extraps,
                                     knots_x=[None] * S, knots_y=[None] * S)
        x = f32(randn32(3, S, *lat, seed=192) * 1.5)
        inv_in = f32(torch.tanh(randn32(3, S, *lat, seed=196)) * 1.8)
        _record_flow(tag, cpl, x, res, seed=193, inv_in=inv_in)
    # controlled couplings: the generator hands out a recorded control field
    for tag, cls, P, kw in [("cntr_affine", CntrAffineCoupling_, 2, {}),
                            ("cntr_shift", CntrShiftCoupling_, 1, {}),
                            ("cntr_rqs", CntrRQSplineCoupling_, 3 * K - 2,
                             dict(xlim=(-4, 4), ylim=(-4, 4), extrap=lin))]:
        torch.manual_seed(201)
        nets = [ConvAct(1, P, 3, hidden_sizes=[4], acts=('tanh', None)) for _ in range(3)]
        for n in nets:
            round_params(n)
        control = randn32(3, *lat, seed=202)
        cpl = cls(nets, mask=mask, control_generator=lambda B, c=control: c[:B], **kw)
        x = f32(randn32(3, *lat, seed=203) * 1.5)
        inv_in = f32(torch.tanh(randn32(3, *lat, seed=206)) * 1.8)
        cpl(x.detach())                  # a forward call sets the control used by backward
        _record_flow(tag, cpl, x, res, seed=204, inv_in=inv_in)
        res[f"{tag}_control"] = npy(control)
    save("rank4_couplings", **res)


def gen_blocked_mcmc():
    """BlockedMCMCSampler (mcmc.py:132-219) on a 4x4 lattice with an affine flow: the block proposals
    drawn by the reference are recorded so that the same chain can be replayed elsewhere; decisions
    use np.random (seed 9) as in the reference."""
    torch.manual_seed(211)
    lat = (4, 4)
    nets = [ConvAct(1, 2, 3, hidden_sizes=[4], acts=('tanh', None)) for _ in range(2)]
    for n in nets:
        round_params(n)
    net_ = ModuleList_([AffineCoupling_(nets, mask=EvenOddMask(shape=lat))])
    prior = NormalPrior(shape=lat)
    model = Model(net_=net_, prior=prior, action=ScalarPhi4Action(**ACTION))
    sampler = BlockedMCMCSampler(model)
    x0 = randn32(1, *lat, seed=212)
    draws = []
    g = torch.Generator('cpu').manual_seed(213)

    def fake_sample(batch_size=1):
        return x0.clone()
    prior.sample = fake_sample
    orig_setup = prior.setup_blockupdater

    def setup(block_len):
        orig_setup(block_len)

        def chopped_sample(batch_size=1):
            d = torch.randn(batch_size, block_len, generator=g, dtype=torch.float32).double()
            draws.append(npy(d))
            return d
        prior.blockupdater.chopped_prior.sample = chopped_sample
    prior.setup_blockupdater = setup
    res = dict(x0=npy(x0), lat=np.array(lat))
    np.random.seed(9)
    for call, (B, nb) in enumerate([(6, 4), (5, 4), (4, 2)]):
        cfgs, logq, logp = sampler.sample__(batch_size=B, n_blocks=nb, bookkeeping=True)
        res.update({f"call{call}_cfgs": npy(cfgs), f"call{call}_logq": npy(logq), f"call{call}_logp": npy(logp),
                    f"call{call}_accept_seq": sampler.history.accept_seq[-1],
                    f"call{call}_accept_rate": np.array(sampler.history.accept_rate[-1]),
                    f"call{call}_shape": np.array([B, nb])})
    res["draws"] = np.concatenate([d.ravel() for d in draws])
    res["draw_lens"] = np.array([d.size for d in draws])
    for key, val in conv_layers(nets[0]).items():
        res[f"step0_{key}"] = val
    for key, val in conv_layers(nets[1]).items():
        res[f"step1_{key}"] = val
    save("blocked_mcmc", **res)


if __name__ == "__main__":
    only = set(sys.argv[1:])           # e.g. `make_golden.py psd model_psd_affine`

    def wanted(name):
        return not only or name in only
    if wanted("masks"):
        gen_masks()
    if wanted("action"):
        gen_action()
    if wanted("prior"):
        gen_prior()
    if wanted("spline"):
        gen_spline()
    if wanted("spline_extra"):
        gen_spline_extra()
    if wanted("rqs_kernel_only"):
        gen_rqs_kernel_only()
    if wanted("affine_kernel_only"):
        gen_affine_kernel_only()
    if wanted("distconv"):
        gen_distconv()
    if wanted("mcmc"):
        gen_mcmc()
    if wanted("conv"):
        gen_conv()
    if wanted("model_zero_dim"):
        gen_model_zero_dim()
    # whole coupling stacks: config-2 style (affine), config-3 style (rqs), 1-D shift,
    # 3-D mixed (config-4 style), 4-D with Conv4d (config-5 style)
    if wanted("cpl_affine_2d"):
        gen_coupling("cpl_affine_2d", (8, 8), 4, [("affine", 4)], seed=10)
    if wanted("cpl_rqs_2d"):
        gen_coupling("cpl_rqs_2d", (8, 8), 3, [("rqs", 4)], seed=20)
    if wanted("cpl_shift_1d"):
        gen_coupling("cpl_shift_1d", (10,), 4, [("shift", 2)], seed=30, hidden=(4,), bias=True)
    if wanted("cpl_mixed_3d"):
        gen_coupling("cpl_mixed_3d", (4, 4, 4), 2, [("affine", 2), ("rqs", 2)], seed=40, hidden=(4, 4))
    if wanted("cpl_mixed_4d"):
        gen_coupling("cpl_mixed_4d", (4, 4, 4, 4), 2, [("affine", 2), ("rqs", 2)], seed=50, hidden=(4,),
                     bias=True)
    # multi-strip 2-D geometries (the fused tensor-core forward walks several strips per sample; the
    # training forward + tensor-core weight gradient + checkerboard data gradient are pinned to the
    # reference's autograd here, not to the layer-wise path)
    if wanted("cpl_rqs_2d_32"):
        gen_coupling("cpl_rqs_2d_32", (32, 32), 2, [("rqs", 2)], seed=60, per_step=False)
    if wanted("cpl_mixed_2d_40x24"):
        gen_coupling("cpl_mixed_2d_40x24", (40, 24), 2, [("affine", 2), ("rqs", 2)], seed=70, per_step=False)
    # 3-D / 4-D stacks with the [8, 8] conditioner: the N-D tensor-core training forward and the tiled N-D
    # weight-gradient kernels against reference autograd
    if wanted("cpl_mixed_3d_h8"):
        gen_coupling("cpl_mixed_3d_h8", (4, 6, 8), 2, [("affine", 1), ("rqs", 2)], seed=80, per_step=False)
    if wanted("cpl_mixed_4d_h8"):
        gen_coupling("cpl_mixed_4d_h8", (4, 4, 4, 4), 2, [("affine", 1), ("rqs", 1)], seed=90, per_step=False)
    if wanted("psd"):
        gen_psd()
    if wanted("model_psd_affine"):
        gen_model_psd_affine()
    if wanted("snapshot"):
        gen_snapshot()
    if wanted("rank4_couplings"):
        gen_rank4_couplings()
    if wanted("blocked_mcmc"):
        gen_blocked_mcmc()
