"""Pin the oracle's restatement of MultiRQSplineCoupling_, the controlled couplings and
BlockedMCMCSampler against the reference's recorded outputs (tests/golden/rank4_couplings.npz,
blocked_mcmc.npz).  CPU only."""

import numpy as np
import pytest

from oracle import nf_oracle as O
from conftest import load_golden

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)
LIN = dict(left='linear', right='linear')
MULTI = {"multi_uniform": dict(xlims=[(-3, 3)] * 2, ylims=[(-3, 3)] * 2, extraps=[LIN, LIN]),
         "multi_mixed": dict(xlims=[(-3, 3), (-2, 2.5), (-4, 4)], ylims=[(-3, 3), (-2, 2.5), (-4, 4)],
                             extraps=[LIN, LIN, {}])}
CNTR = {"cntr_affine": ('affine', {}), "cntr_shift": ('shift', {}),
        "cntr_rqs": ('rqs', dict(xlim=(-4, 4), ylim=(-4, 4), extrap=LIN))}


def convact_layers(g, tag, k, prefix="nets.{k}."):
    """[(w, b)] of a ConvAct(.., hidden_sizes=[4], acts=('tanh', None)) stored under `<tag>_w_<name>`."""
    pre = f"{tag}_w_" + prefix.format(k=k)
    return [(g[pre + f"{i}.weight"], g[pre + f"{i}.bias"]) for i in (0, 2)]


def multi_flow(g, tag):
    mask = O.evenodd_mask((4, 6))
    steps = [O.make_multi_rqs_step(convact_layers(g, tag, k, "nets.{k}.net."), ('tanh', None), mask, **MULTI[tag])
             for k in range(3)]
    return lambda x, log0, inverse=False: O.coupling_forward(x, log0, mask, steps, inverse=inverse)


def cntr_flow(g, tag):
    mask = O.evenodd_mask((4, 6))
    kind, kw = CNTR[tag]
    steps = [O.make_convact_step(kind, convact_layers(g, tag, k), ('tanh', None), mask, **kw) for k in range(3)]
    control = g[f"{tag}_control"]
    return lambda x, log0, inverse=False: O.cntr_coupling_forward(x, control, log0, mask, steps, inverse=inverse)


@pytest.mark.parametrize("tag", sorted(MULTI) + sorted(CNTR))
def test_rank4_couplings_golden(tag):
    g = load_golden("rank4_couplings")
    flow = multi_flow(g, tag) if tag in MULTI else cntr_flow(g, tag)
    y, logJ = flow(g[f"{tag}_x"], 0.0)
    np.testing.assert_allclose(y, g[f"{tag}_y"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(logJ * np.ones(len(y)), g[f"{tag}_logJ"], rtol=1e-10, atol=1e-10)
    xb, lb = flow(g[f"{tag}_inv_in"], g[f"{tag}_c"], inverse=True)
    np.testing.assert_allclose(xb, g[f"{tag}_inv_x"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(lb, g[f"{tag}_inv_log"], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("tag", ["multi_mixed", "cntr_rqs"])
def test_rank4_input_gradient_by_central_differences(tag):
    g = load_golden("rank4_couplings")
    flow = multi_flow(g, tag) if tag in MULTI else cntr_flow(g, tag)
    x, r, c = g[f"{tag}_x"], g[f"{tag}_r"], g[f"{tag}_c"]

    def scalar(v):
        y, logJ = flow(v, 0.0)
        return (y * r).sum() + (logJ * c).sum()
    rng = np.random.RandomState(4)
    for _ in range(8):
        idx = tuple(rng.randint(0, n) for n in x.shape)
        xp, xm = x.copy(), x.copy()
        xp[idx] += 1e-6
        xm[idx] -= 1e-6
        np.testing.assert_allclose((scalar(xp) - scalar(xm)) / 2e-6, g[f"{tag}_gx"][idx], rtol=1e-5, atol=1e-7)


def blocked_setup(g):
    lat = tuple(int(v) for v in g["lat"])
    mask = O.evenodd_mask(lat)
    steps = [O.make_convact_step('affine', [(g[f"step{k}_w0"], g[f"step{k}_b0"]), (g[f"step{k}_w1"], g[f"step{k}_b1"])],
                                 ('tanh', None), mask) for k in range(2)]

    def evaluate(x):
        y, logJ = O.coupling_forward(x, 0.0, mask, steps)
        return y, O.normal_log_prob(x) - logJ, -O.phi4_action(y, **ACTION)

    def inverse(y):
        return O.coupling_forward(y, 0.0, mask, steps, inverse=True)[0]
    return evaluate, inverse


def test_blocked_mcmc_chain_golden():
    """Replay the reference's chain: same start, same recorded block proposals, same np.random stream."""
    g = load_golden("blocked_mcmc")
    evaluate, inverse = blocked_setup(g)
    lens = g["draw_lens"]
    flat = g["draws"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    proposals = iter([flat[offs[i]:offs[i + 1]].reshape(1, -1) for i in range(len(lens))])
    np.random.seed(9)
    x = g["x0"].copy()
    ref = None
    for call in range(3):
        B, nb = (int(v) for v in g[f"call{call}_shape"])
        if call > 0:
            x = inverse(last[None])
        unif = iter([np.log(np.random.rand(nb)) for _ in range(B)])
        cfgs, logq, logp, acc, ref, x = O.blocked_mcmc(x, evaluate, proposals, unif, B, nb, ref)
        np.testing.assert_array_equal(acc.ravel(), g[f"call{call}_accept_seq"])
        np.testing.assert_allclose(cfgs, g[f"call{call}_cfgs"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(logq, g[f"call{call}_logq"], rtol=1e-10)
        np.testing.assert_allclose(logp, g[f"call{call}_logp"], rtol=1e-10)
        np.testing.assert_allclose(acc.mean(), g[f"call{call}_accept_rate"])
        last = cfgs[-1]
    assert next(proposals, None) is None          # every recorded draw was consumed
