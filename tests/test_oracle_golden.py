"""Pin the numpy oracle against the reference's own outputs (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference) and against the
known-answer values of SURVEY.md section 4.  CPU only."""

import numpy as np
import pytest

from oracle import nf_oracle as O
from conftest import load_golden

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)
TIGHT = dict(rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------- masks (bit-exact)
def test_masks_bit_exact():
    g = load_golden("masks")
    n = 0
    for key in g.files:
        if key.endswith("_meta"):
            tag = key[:-5]
            meta = g[key]
            parity, mu, shape = int(meta[0]), int(meta[1]), tuple(int(v) for v in meta[2:])
            if tag.startswith("eo_"):
                m = O.evenodd_mask(shape, parity, None if mu < 0 else mu)
                assert np.array_equal(1 - m, g[tag + "_cmask"])
            else:
                m = O.alongaxes_mask(shape, parity, mu)
            assert m.dtype == np.uint8 and np.array_equal(m, g[tag + "_mask"])
            n += 1
    assert n == 13


def test_mask_known_answers():
    m = O.evenodd_mask((4, 4))
    assert m[0].tolist() == [1, 0, 1, 0] and m.sum() == 8
    assert O.evenodd_mask((4, 4), parity=1)[0].tolist() == [0, 1, 0, 1]


# ---------------------------------------------------------------- action
def test_action_golden():
    g = load_golden("action")
    for i in range(9):
        cfgs = g[f"a{i}_cfgs"]
        np.testing.assert_allclose(O.phi4_action(cfgs, **ACTION), g[f"a{i}_S"], **TIGHT)
        np.testing.assert_allclose(O.phi4_action_density(cfgs, **ACTION), g[f"a{i}_density"], **TIGHT)
        gw = g[f"a{i}_gS"].reshape((-1,) + (1,) * (cfgs.ndim - 1))
        np.testing.assert_allclose(O.phi4_action_grad(cfgs, **ACTION) * gw, g[f"a{i}_gcfgs"], **TIGHT)
    np.testing.assert_allclose(O.phi4_action(g["zd_cfgs"], kappa=0, m_sq=-1.2, lambd=0.5), g["zd_S"], **TIGHT)
    np.testing.assert_allclose(O.phi4_action(g["lat_cfgs"], kappa=0.5, m_sq=0.3, lambd=0.2, a=0.5),
                               g["lat_S"], **TIGHT)


@pytest.mark.parametrize("shape,expect", [((16, 16), -123.84), ((64, 64), -1981.44),
                                          ((8, 8, 8), -247.68), ((4, 4, 4, 4), -123.84)])
def test_action_constant_field(shape, expect):
    S = O.phi4_action(np.full((2,) + shape, 1.5), **ACTION)
    np.testing.assert_allclose(S, expect, rtol=1e-12)


def test_action_short_axes():
    # L=1: the roll term is a self-interaction; L=2: each bond counted twice
    assert np.isclose(O.phi4_action(np.array([[3.0]]), kappa=1, m_sq=1, lambd=0)[0], 4.5)
    assert np.isclose(O.phi4_action(np.array([[1.0, 2.0]]), kappa=1, m_sq=1, lambd=0)[0], 3.5)


def test_action_density_sums_to_action():
    x = np.random.RandomState(0).randn(3, 6, 6)
    np.testing.assert_allclose(O.phi4_action_density(x, **ACTION).sum(axis=(1, 2)),
                               O.phi4_action(x, **ACTION), rtol=1e-12)


# ---------------------------------------------------------------- prior
def test_prior_golden():
    g = load_golden("prior")
    np.testing.assert_allclose(O.normal_log_prob(g["std_x"]), g["std_logr"], **TIGHT)
    np.testing.assert_allclose(O.normal_log_prob(g["gen_x"], g["gen_loc"], g["gen_scale"]),
                               g["gen_logr"], **TIGHT)


# ---------------------------------------------------------------- spline
MODES = {"none": {}, "lin": dict(left='linear', right='linear'), "linleft": dict(left='linear'),
         "linright": dict(right='linear'), "antileft": dict(left='anti'),
         "antiright": dict(right='anti'), "antilin": dict(left='anti', right='linear')}


@pytest.mark.parametrize("mode", sorted(MODES))
def test_spline_golden(mode):
    g = load_golden("spline")
    sp = O.RQSpline(g["kx"], g["ky"], g["kd"], axis=1, extrap=MODES[mode])
    assert sp.kx.shape[1] == int(g[f"{mode}_nknots"])
    y, gr = sp.forward(g["x"])
    np.testing.assert_allclose(y, g[f"{mode}_y"], **TIGHT)
    np.testing.assert_allclose(gr, g[f"{mode}_g"], **TIGHT)
    xi, gi = sp.backward(g[f"{mode}_y"])
    # the reference's own inverse (NaN/garbage included where it is ill-conditioned)
    np.testing.assert_allclose(xi, g[f"{mode}_xinv"], rtol=1e-9, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(gi, g[f"{mode}_ginv"], rtol=1e-9, atol=1e-9, equal_nan=True)


def test_spline_shared_knots_smooth_anti():
    g = load_golden("spline")
    sp = O.RQSpline(g["s1_kx"], g["s1_ky"], None, extrap=dict(left='anti'))
    np.testing.assert_allclose(sp.kx, g["s1_kxaug"], **TIGHT)
    np.testing.assert_allclose(sp.kd, g["s1_kd"], **TIGHT)
    y, gr = sp.forward(g["s1_x"])
    np.testing.assert_allclose(y, g["s1_y"], **TIGHT)
    np.testing.assert_allclose(gr, g["s1_g"], **TIGHT)


def test_spline_roundtrip_in_range_and_stable_root():
    g = load_golden("spline")
    for mode in ("none", "lin", "antilin"):
        sp = O.RQSpline(g["kx"], g["ky"], g["kd"], axis=1, extrap=MODES[mode], stable_inverse=True)
        x = g["x"]
        y, gr = sp.forward(x)
        xi, gi = sp.backward(y)
        if mode == "none":   # extrapolated first/last segments are not monotone in general
            inside = (x > g["kx"][:, :1]) & (x < g["kx"][:, -1:])
        else:
            inside = np.ones_like(x, dtype=bool)
        np.testing.assert_allclose(xi[inside], x[inside], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose((gi * gr)[inside], 1.0, rtol=1e-8)


# ---------------------------------------------------------------- kernel-only coupling steps
def test_affine_kernel_golden():
    g = load_golden("affine_kernel")
    mask = g["mask"]
    for parity in (0, 1):
        xa = O.mask_purify(mask, g["x"], parity)
        fx, logJ = O.affine_atomic(xa, g["out"], mask, parity, 0.0)
        np.testing.assert_allclose(fx, g[f"p{parity}_fx"], **TIGHT)
        np.testing.assert_allclose(logJ, g[f"p{parity}_logJ"], **TIGHT)
        xi, li = O.affine_atomic(fx, g["out"], mask, parity, 0.0, inverse=True)
        np.testing.assert_allclose(xi, g[f"p{parity}_xinv"], **TIGHT)
        np.testing.assert_allclose(li, g[f"p{parity}_loginv"], **TIGHT)


@pytest.mark.parametrize("tag,xlim,ylim,extrap", [
    ("lin", (-5, 5), (-4, 6), dict(left='linear', right='linear')),
    ("none", (-5, 5), (-4, 6), {}),
    ("mixed", (-5, 5), (-4, 6), dict(left='linear'))])
def test_rqs_kernel_golden(tag, xlim, ylim, extrap):
    g = load_golden("rqs_kernel")
    mask = g[f"{tag}_mask"]
    for parity in (0, 1):
        xa = O.mask_purify(mask, g[f"{tag}_x"], parity)
        fx, logJ = O.rqs_atomic(xa, g[f"{tag}_out"], mask, parity, 0.0, xlim=xlim, ylim=ylim, extrap=extrap)
        np.testing.assert_allclose(fx, g[f"{tag}_p{parity}_fx"], **TIGHT)
        np.testing.assert_allclose(logJ, g[f"{tag}_p{parity}_logJ"], **TIGHT)


def test_rqs_kernel_gradient_via_complex_step():
    """The oracle's complex-step derivative reproduces reference autograd."""
    g = load_golden("rqs_kernel")
    tag, parity = "lin", 1
    mask, r, c = g[f"{tag}_mask"], g[f"{tag}_r"], g[f"{tag}_c"]
    kw = dict(xlim=(-5, 5), ylim=(-4, 6), extrap=dict(left='linear', right='linear'))

    def scalar(xa, out):
        fx, logJ = O.rqs_atomic(xa, out, mask, parity, 0.0, **kw)
        return (fx * r).sum() + (logJ * c).sum()

    rs = np.random.RandomState(3)
    xa = O.mask_purify(mask, g[f"{tag}_x"], parity)
    vx, vo = rs.randn(*xa.shape), rs.randn(*g[f"{tag}_out"].shape)
    dd = O.directional_derivative(scalar, (xa, g[f"{tag}_out"]), (vx, vo))
    expect = (g[f"{tag}_p{parity}_gx"] * vx).sum() + (g[f"{tag}_p{parity}_gout"] * vo).sum()
    np.testing.assert_allclose(dd, expect, rtol=1e-10)


# ---------------------------------------------------------------- conv stacks
def _layers(g, prefix, n):
    layers = []
    for i in range(n):
        b = g[f"{prefix}_b{i}"] if f"{prefix}_b{i}" in g.files else None
        layers.append((g[f"{prefix}_w{i}"], b))
    return layers


@pytest.mark.parametrize("tag", ["d1", "d2", "d3", "d4"])
def test_convact_golden(tag):
    g = load_golden("conv")
    n = len(g[f"{tag}_hidden"]) + 1
    layers = _layers(g, tag, n)
    out = O.convact_forward(g[f"{tag}_x"], layers, ['tanh'] * (n - 1) + [None])
    np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=1e-11, atol=1e-12)
    if tag == "d4":   # stored Conv4d layout -> standard layout
        w = O.conv4d_lower_to_standard(g["d4_wlower0"], g["d4_w0"].shape[0], 3)
        assert np.array_equal(w, g["d4_w0"])


# ---------------------------------------------------------------- whole coupling stacks
def _run_stack(g, inverse=False):
    mask = g["mask"]
    blocks = [s.split(":") for s in g["blocks"]]
    acts = [None if a == 'none' else str(a) for a in g["acts"]]
    n_layers = len(g["hidden"]) + 1
    flows = []
    for bi, (kind, n_steps) in enumerate(blocks):
        steps = []
        for k in range(int(n_steps)):
            layers = _layers(g, f"blk{bi}_step{k}", n_layers)
            kw = dict(xlim=(-5, 5), ylim=(-5, 5), extrap=dict(left='linear', right='linear'),
                      stable_inverse=True) if kind == 'rqs' else {}
            steps.append(O.make_convact_step(kind, layers, acts, mask, **kw))
        flows.append(steps)
    return mask, flows


@pytest.mark.parametrize("name", ["cpl_affine_2d", "cpl_rqs_2d", "cpl_shift_1d", "cpl_mixed_3d", "cpl_mixed_4d",
                                  "cpl_rqs_2d_32", "cpl_mixed_2d_40x24", "cpl_mixed_3d_h8", "cpl_mixed_4d_h8"])
def test_coupling_stack_golden(name):
    g = load_golden(name)
    mask, flows = _run_stack(g)
    x = g["x"]
    log = np.zeros(x.shape[0])
    y = x
    for bi, steps in enumerate(flows):
        y, log = O.coupling_forward(y, log, mask, steps)
        np.testing.assert_allclose(y, g[f"blk{bi}_y"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(log, g[f"blk{bi}_logJ"] + np.zeros_like(log), rtol=1e-10, atol=1e-10)
    S = O.phi4_action(y, **ACTION)
    logr = O.normal_log_prob(x)
    np.testing.assert_allclose(S, g["S"], rtol=1e-10)
    np.testing.assert_allclose(O.kl_loss(logr - log, -S), g["loss"], rtol=1e-10)
    # inverse with the stable root: back to x, residual log ~ 0
    xb, lb = y, log
    for steps in reversed(flows):
        xb, lb = O.coupling_forward(xb, lb, mask, steps, inverse=True)
    np.testing.assert_allclose(xb, x, rtol=1e-8, atol=1e-8)
    # the forward log0 is ADDED to by the inverse: backward(y, log0=logJ) -> ~0
    np.testing.assert_allclose(lb, 0.0, atol=1e-8)


def test_coupling_first_block_steps():
    g = load_golden("cpl_rqs_2d")
    mask = g["mask"]
    for k in range(4):
        p = k % 2
        fx, _ = O.rqs_atomic(g[f"blk0_step{k}_xactive"], g[f"blk0_step{k}_out"], mask, p, 0.0,
                             xlim=(-5, 5), ylim=(-5, 5), extrap=dict(left='linear', right='linear'))
        np.testing.assert_allclose(fx, g[f"blk0_step{k}_fx"], rtol=1e-11, atol=1e-12)


# ---------------------------------------------------------------- DistConvertor_
@pytest.mark.parametrize("tag,symmetric", [("zd_sym", True), ("lat_sym_smooth", True), ("lat_asym", False)])
def test_distconvertor_golden(tag, symmetric):
    g = load_golden("distconv")
    wd = g[f"{tag}_wd"] if f"{tag}_wd" in g.files else None
    weights = (g[f"{tag}_wx"], g[f"{tag}_wy"], wd)
    y, logJ = O.distconvertor(g[f"{tag}_x"], 0.0, weights, symmetric=symmetric)
    np.testing.assert_allclose(y, g[f"{tag}_y"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(logJ, g[f"{tag}_logJ"], rtol=1e-10, atol=1e-10)
    xb, lb = O.distconvertor(g[f"{tag}_y"], g[f"{tag}_logJ"], weights, symmetric=symmetric, inverse=True)
    np.testing.assert_allclose(xb, g[f"{tag}_inv_x"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(lb, g[f"{tag}_inv_log"], atol=1e-7)


def test_distconvertor_weight_grads_via_complex_step():
    g = load_golden("distconv")
    tag = "zd_sym"
    r, c, x = g[f"{tag}_r"], g[f"{tag}_c"], g[f"{tag}_x"]

    def scalar(wx, wy, wd):
        y, logJ = O.distconvertor(x, 0.0, (wx, wy, wd), symmetric=True)
        return (y * r).sum() + (logJ * c).sum()

    args = (g[f"{tag}_wx"], g[f"{tag}_wy"], g[f"{tag}_wd"])
    names = [str(n) for n in g[f"{tag}_param_names"]]
    for j, key in enumerate(("weights_x", "weights_y", "weights_d")):
        ref = g[f"{tag}_grad_" + [n for n in names if n.endswith(key)][0]]
        for i in range(len(ref)):
            v = [np.zeros_like(a) for a in args]
            v[j][i] = 1.0
            np.testing.assert_allclose(O.directional_derivative(scalar, args, v), ref[i], rtol=1e-9, atol=1e-11)


def test_model_zero_dim_golden():
    g = load_golden("model_zero_dim")
    weights = tuple(g[f"w_1.weights_{k}"] for k in "xyd")
    flow = lambda x, log0: O.distconvertor(x, log0, weights, symmetric=True)
    y, logq, logp = O.posterior_sample__(g["x"], flow, dict(kappa=0, m_sq=-1.2, lambd=0.5))
    np.testing.assert_allclose(y, g["y"], rtol=1e-10)
    np.testing.assert_allclose(logq, g["logq"], rtol=1e-10)
    np.testing.assert_allclose(logp, g["logp"], rtol=1e-10)
    np.testing.assert_allclose(O.kl_loss(logq, logp), g["loss"], rtol=1e-10)


# ---------------------------------------------------------------- Metropolis
def test_metropolis_golden():
    g = load_golden("mcmc")
    st = O.metropolis_accept_status(g["logqp"], g["u_first"])
    assert st[0] and np.array_equal(st, g["status_noref"])
    assert np.array_equal(O.metropolis_accept_indices(st), g["ind_noref"])
    st = O.metropolis_accept_status(g["logqp"], g["u_second"], float(g["ref"]))
    assert np.array_equal(st, g["status_ref"])
    assert np.array_equal(O.metropolis_accept_indices(st), g["ind_ref"])


def test_mcmc_accept_reject_chain_golden():
    g = load_golden("mcmc")
    ref = dict(sample=None, logq=None, logp=None, logqp=None)
    for call in range(2):
        y, lq, lp, acc, idx = O.mcmc_accept_reject(g[f"c{call}_y"], g[f"c{call}_logq"], g[f"c{call}_logp"],
                                                   g[f"c{call}_u"], ref)
        assert np.array_equal(y, g[f"c{call}_yo"])
        assert np.array_equal(lq, g[f"c{call}_logqo"]) and np.array_equal(lp, g[f"c{call}_logpo"])
        assert np.isclose(acc.mean(), float(g[f"c{call}_accept_rate"]))


def test_zero_dim_logz_known_answer():
    """Analytic log Z of the config-1 target: 1.112773."""
    xs = np.linspace(-6, 6, 200001)
    S = O.phi4_action(xs[:, None], kappa=0, m_sq=-1.2, lambd=0.5)
    logz = np.log(np.trapezoid(np.exp(-S), xs))
    assert abs(logz - 1.112773) < 1e-6


def test_torch_cpu_port_matches_numpy_oracle():
    """oracle/torch_port.py (the CPU baseline bench.py times) against the numpy oracle on the
    bench workload's operator chain at a small size."""
    import torch
    from oracle import torch_port as T
    rs = np.random.RandomState(3)
    shape, B, K = (8, 12), 5, 10
    sizes = [1, 8, 8, 3 * K - 2]
    nets_np = [[rs.randn(sizes[i + 1], sizes[i], 3, 3) / np.sqrt(9 * sizes[i]) for i in range(3)] for _ in range(3)]
    x = rs.randn(B, *shape) * 1.5
    mask = O.evenodd_mask(shape)
    steps = [O.make_convact_step('rqs', [(w, None) for w in ws], ['tanh', 'tanh', None], mask, xlim=(-5, 5),
                                 ylim=(-5, 5), extrap=dict(left='linear', right='linear')) for ws in nets_np]
    action = dict(kappa=0.67, m_sq=-2.68, lambd=0.5)
    y_np, logq_np, logp_np = O.posterior_sample__(x, lambda v, l0: O.coupling_forward(v, l0, mask, steps), action)
    nets_t = [[torch.tensor(w, dtype=torch.float64, device='cpu') for w in ws] for ws in nets_np]
    y_t, logq_t, logp_t = T.posterior_sample__(torch.tensor(x, dtype=torch.float64, device='cpu'), nets_t,
                                               T.evenodd_mask(shape), (-5.0, 5.0), (-5.0, 5.0), action)
    assert np.array_equal(T.evenodd_mask(shape).numpy(), mask)
    np.testing.assert_allclose(y_t.numpy(), y_np, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(logq_t.numpy(), logq_np, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(logp_t.numpy(), logp_np, rtol=1e-12, atol=1e-11)


# ------------------------------------------------------------------ periodic extrapolation, fixed knots
@pytest.mark.parametrize("name,extrap", [("perleft", dict(left='periodic')), ("perright", dict(right='periodic')),
                                         ("perboth", dict(left='periodic', right='periodic')),
                                         ("perleft_antiright", dict(left='periodic', right='anti'))])
def test_periodic_extrapolation_golden(name, extrap):
    """AugmentKnots 'periodic' (spline.py:502-508, 518-524): even mirror image, derivative sign flipped."""
    g = load_golden("spline_extra")
    sp = O.RQSpline(g["per_kx"], g["per_ky"], g["per_kd"], extrap=extrap)
    y, gr = sp.forward(g["per_x"])
    np.testing.assert_allclose(y, g[f"{name}_y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(gr, g[f"{name}_g"], rtol=1e-10, atol=1e-12)
    assert (g[f"{name}_g"] < 0).any()          # the fixture does visit the mirrored range


def _fixed_knot_steps(g, tag):
    mask = O.evenodd_mask(tuple(int(v) for v in g["fix_shape"]))
    kw = {"fixx": dict(knots_x=g["fix_kx"], extrap=dict(left='linear', right='linear')),
          "fixy": dict(knots_y=g["fix_ky"], extrap={}),
          "fixxy": dict(knots_x=g["fix_kx"], knots_y=g["fix_ky"], extrap={})}[tag]
    steps = []
    for k in range(2):
        layers = [(g[f"{tag}_step{k}_w{i}"], g[f"{tag}_step{k}_b{i}"]) for i in range(2)]
        steps.append(O.make_convact_step('rqs', layers, ['tanh', None], mask, xlim=(-3, 3), ylim=(-2.5, 2.5),
                                         stable_inverse=True, **kw))
    return mask, steps


@pytest.mark.parametrize("tag", ["fixx", "fixy", "fixxy"])
def test_fixed_knots_coupling_golden(tag):
    """RQSplineCoupling_(knots_x=..., knots_y=...) (couplings_.py:246-258)."""
    g = load_golden("spline_extra")
    mask, steps = _fixed_knot_steps(g, tag)
    x = g[f"{tag}_x"]
    y, log = O.coupling_forward(x, np.zeros(x.shape[0]), mask, steps)
    np.testing.assert_allclose(y, g[f"{tag}_y"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(log, g[f"{tag}_logJ"], rtol=1e-10, atol=1e-10)
    xb, lb = O.coupling_forward(y, log, mask, steps, inverse=True)
    np.testing.assert_allclose(xb, x, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(lb, 0.0, atol=1e-8)
