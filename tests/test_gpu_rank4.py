"""GPU parity of SURVEY section 8f rank 4: MultiRQSplineCoupling_, the controlled couplings and
BlockedMCMCSampler, against the reference's recorded outputs and the oracle.  The accept / reject
decisions of the blocked chain are bit-exact given the same proposals and the same np.random stream."""

import numpy as np
import pytest
import torch

from conftest import load_golden
from test_oracle_rank4 import MULTI, CNTR, blocked_setup

import normflow__b200 as nf  # noqa: F401
from normflow__b200 import Model
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import (ModuleList_, ConvAct, AffineCoupling_, MultiRQSplineCoupling_, CntrAffineCoupling_,
                               CntrShiftCoupling_, CntrRQSplineCoupling_, CntrMultiRQSplineCoupling_,
                               DirectCntrCoupling_, Coupling_)
from normflow__b200.prior import NormalPrior
from test_gpu_parity import cu, close, close_grad, DEV, ACTION

pytestmark = pytest.mark.gpu
LAT = (4, 6)
K = 5


class SqueezeChannel(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return self.net(x.squeeze(1))


def _load(net_, g, tag):
    names = [str(n) for n in g[f"{tag}_param_names"]]
    assert [n for n, _ in net_.named_parameters()] == names
    with torch.no_grad():
        for n, p in net_.named_parameters():
            p.copy_(cu(g[f"{tag}_w_{n}"]))
    return names


def _build(tag, g):
    mask = EvenOddMask(shape=LAT)
    if tag in MULTI:
        S = len(MULTI[tag]['xlims'])
        nets = [SqueezeChannel(ConvAct(S, S * (3 * K - 2), 3, hidden_sizes=[4], acts=('tanh', None))) for _ in range(3)]
        return MultiRQSplineCoupling_(nets, mask=mask, **MULTI[tag])
    kind, kw = CNTR[tag]
    cls, P = {'affine': (CntrAffineCoupling_, 2), 'shift': (CntrShiftCoupling_, 1),
              'rqs': (CntrRQSplineCoupling_, 3 * K - 2)}[kind]
    nets = [ConvAct(1, P, 3, hidden_sizes=[4], acts=('tanh', None)) for _ in range(3)]
    control = cu(g[f"{tag}_control"])
    return cls(nets, mask=mask, control_generator=lambda B: control[:B], **kw)


@pytest.mark.parametrize("tag", sorted(MULTI) + sorted(CNTR))
def test_rank4_couplings_golden(tag):
    g = load_golden("rank4_couplings")
    cpl = _build(tag, g)
    names = _load(cpl, g, tag)
    x = cu(g[f"{tag}_x"]).requires_grad_(True)
    y, logJ = cpl(x)
    close(y, g[f"{tag}_y"])
    logJ_full = logJ if torch.is_tensor(logJ) else torch.zeros(x.shape[0], device=DEV) + logJ
    close(logJ_full, g[f"{tag}_logJ"])
    L = (y * cu(g[f"{tag}_r"])).sum() + (logJ_full * cu(g[f"{tag}_c"])).sum()
    grads = torch.autograd.grad(L, [x] + list(cpl.parameters()), allow_unused=True)
    close_grad(grads[0], g[f"{tag}_gx"])
    for n, gr in zip(names, grads[1:]):
        ref = g[f"{tag}_grad_{n}"]
        if gr is None:
            assert np.all(ref == 0), n
        else:
            close_grad(gr, ref, tol=2e-5)
    with torch.no_grad():
        xb, lb = cpl.backward(cu(g[f"{tag}_inv_in"]), log0=cu(g[f"{tag}_c"]))
    close(xb, g[f"{tag}_inv_x"], tol=2e-5)
    close(lb, g[f"{tag}_inv_log"], tol=2e-5)


def test_multi_rqs_uniform_batched_launch_equals_per_component_launches():
    """The (B*S) single-launch path and the S-launch path are the same arithmetic."""
    g = load_golden("rank4_couplings")
    cpl = _build("multi_uniform", g)
    _load(cpl, g, "multi_uniform")
    x = cu(g["multi_uniform_x"])
    with torch.no_grad():
        y1, l1 = cpl(x)
        cpl._uniform = lambda: False
        y2, l2 = cpl(x)
    assert torch.equal(y1, y2)
    close(l1, l2.cpu().numpy(), tol=1e-6)


def test_multi_rqs_atomic_api_and_errors():
    g = load_golden("rank4_couplings")
    cpl = _build("multi_mixed", g)
    _load(cpl, g, "multi_mixed")
    x = cu(g["multi_mixed_x"])
    with torch.no_grad():
        y_ref, l_ref = cpl(x)
        y_gen, l_gen = Coupling_.forward(cpl, x)            # split -> atomic_forward -> cat
    close(y_gen, y_ref.cpu().numpy(), tol=1e-6)
    close(l_gen, l_ref.cpu().numpy(), tol=1e-6)
    with pytest.raises(ValueError):
        cpl(x[:, :2])
    with pytest.raises(NotImplementedError):
        MultiRQSplineCoupling_([], mask=EvenOddMask(shape=LAT), knots_x=[torch.zeros(3), None])
    parts = cpl.preprocess(x)
    assert len(parts) == 3 and torch.equal(cpl.postprocess(parts), x)


def test_direct_cntr_coupling_tuple_protocol_and_multi():
    """DirectCntrCoupling_ takes and returns (x, control); CntrMultiRQSplineCoupling_ composes both."""
    g = load_golden("rank4_couplings")
    cpl = _build("cntr_affine", g)
    _load(cpl, g, "cntr_affine")
    x, control = cu(g["cntr_affine_x"]), cu(g["cntr_affine_control"])
    with torch.no_grad():
        (y, c_out), logJ = DirectCntrCoupling_.forward(cpl, (x, control))
        assert c_out is control
        close(y, g["cntr_affine_y"])
        (xb, _), lb = DirectCntrCoupling_.backward(cpl, (y, control), logJ)
        close(xb, g["cntr_affine_x"], tol=2e-5)
        assert float(lb.abs().max()) < 1e-4
    mask = EvenOddMask(shape=LAT)
    torch.manual_seed(3)
    nets = [SqueezeChannel(ConvAct(2, 2 * (3 * K - 2), 3, hidden_sizes=[4], acts=('tanh', None))) for _ in range(2)]
    ctrl = torch.randn(4, 2, *LAT, device=DEV)
    multi = CntrMultiRQSplineCoupling_(nets, mask=mask, control_generator=lambda B: ctrl[:B],
                                       xlims=[(-4, 4)] * 2, ylims=[(-4, 4)] * 2)
    with torch.no_grad():
        x = torch.randn(4, 2, *LAT, device=DEV)
        y, logJ = multi(x)
        xb, lb = multi.backward(y, logJ)
    assert float((xb - x).abs().max()) < 2e-5 and float(lb.abs().max()) < 1e-4


# ------------------------------------------------------------------ blocked MCMC
def _blocked_model(g):
    lat = tuple(int(v) for v in g["lat"])
    nets = [ConvAct(1, 2, 3, hidden_sizes=[4], acts=('tanh', None)) for _ in range(2)]
    with torch.no_grad():
        for k, net in enumerate(nets):
            net[0].weight.copy_(cu(g[f"step{k}_w0"]))
            net[0].bias.copy_(cu(g[f"step{k}_b0"]))
            net[2].weight.copy_(cu(g[f"step{k}_w1"]))
            net[2].bias.copy_(cu(g[f"step{k}_b1"]))
    net_ = ModuleList_([AffineCoupling_(nets, mask=EvenOddMask(shape=lat))])
    model = Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(**ACTION))
    model.device_handler.to(DEV)
    return model


def test_blocked_mcmc_replays_the_reference_chain():
    g = load_golden("blocked_mcmc")
    model = _blocked_model(g)
    prior, sampler = model.prior, model.blocked_mcmc
    lens, flat = g["draw_lens"], g["draws"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    draws = iter([cu(flat[offs[i]:offs[i + 1]].reshape(1, -1)) for i in range(len(lens))])
    x0 = cu(g["x0"])
    prior.sample = lambda batch_size=1: x0.clone()
    orig_setup = prior.setup_blockupdater

    def setup(block_len):
        orig_setup(block_len)
        prior.blockupdater.chopped_prior.sample = lambda batch_size=1: next(draws)
    prior.setup_blockupdater = setup
    np.random.seed(9)
    for call in range(3):
        B, nb = (int(v) for v in g[f"call{call}_shape"])
        cfgs, logq, logp = sampler.sample__(batch_size=B, n_blocks=nb, bookkeeping=True)
        assert np.array_equal(sampler.history.accept_seq[-1], g[f"call{call}_accept_seq"])     # bit-exact decisions
        close(cfgs, g[f"call{call}_cfgs"])
        close(logq, g[f"call{call}_logq"])
        close(logp, g[f"call{call}_logp"])
        assert abs(sampler.history.accept_rate[-1] - float(g[f"call{call}_accept_rate"])) < 1e-12
    assert next(draws, None) is None


def test_blocked_mcmc_with_its_own_generator():
    """End to end with the package's Philox block proposals: the chain moves, rejected blocks are
    restored exactly, the block stream never repeats the prior's own draws, n_blocks must divide nvar."""
    g = load_golden("blocked_mcmc")
    model = _blocked_model(g)
    torch.manual_seed(17)
    np.random.seed(17)
    sampler = model.blocked_mcmc
    first = model.prior.sample(1).clone()
    model.prior.setup_blockupdater(4)
    blk = model.prior.blockupdater.chopped_prior.sample(1)
    assert not torch.equal(blk.reshape(-1), first.reshape(-1)[:4])
    cfgs, logq, logp = sampler.sample__(batch_size=24, n_blocks=4, bookkeeping=True)
    assert cfgs.shape == (24, 4, 4) and torch.isfinite(cfgs).all()
    acc = sampler.history.accept_seq[-1].reshape(24, 4)
    assert 0.05 < acc.mean() <= 1.0
    y, lq, lp = cfgs.cpu().numpy(), logq.cpu().numpy(), logp.cpu().numpy()
    for i in range(1, 24):
        if not acc[i].any():            # a fully rejected sweep repeats the previous configuration
            assert np.array_equal(y[i], y[i - 1]) and lq[i] == lq[i - 1] and lp[i] == lp[i - 1]
    # recorded logq / logp are those of the recorded configurations (oracle on the same y)
    evaluate, inverse = blocked_setup(g)
    x = inverse(y.astype(np.float64))
    _, oq, op = evaluate(x)
    close(logq, oq, tol=5e-5)
    close(logp, op, tol=2e-5)
    with pytest.raises(AssertionError):
        sampler.sample__(batch_size=1, n_blocks=5)
    x = torch.randn(3, 4, 4, device=DEV)
    keep = x.clone()
    model.prior.setup_blockupdater(8)
    model.prior.blockupdater(x, 1)
    assert torch.equal(x.view(3, 2, 8)[:, 0], keep.view(3, 2, 8)[:, 0]) and not torch.equal(x, keep)
    model.prior.blockupdater.restore(x, 1)
    assert torch.equal(x, keep)


def test_blocked_mcmc_block_stream_advances_across_calls(monkeypatch):
    """Consecutive blocked_mcmc.sample__ calls must PROPOSE different block values: the block updater's
    Philox stream keeps its place across calls -- also when the block length changes in between -- (the
    reference draws fresh torch RNG values every time, prior.py:161-178) instead of restarting with every
    setup_blockupdater."""
    from normflow__b200.nn import DistConvertor_
    from normflow__b200.prior.prior import BlockUpdater
    torch.manual_seed(5)
    np.random.seed(5)
    model = Model(net_=DistConvertor_(6, symmetric=True), prior=NormalPrior(shape=(4, 4)),
                  action=ScalarPhi4Action(kappa=0.3, m_sq=-1.2, lambd=0.5))
    model.device_handler.to(DEV)
    seen = []
    orig = BlockUpdater.__call__

    def recording(self, x, block_ind):
        orig(self, x, block_ind)
        seen[-1].append(self._blocks(x)[:, block_ind].clone().flatten())
    monkeypatch.setattr(BlockUpdater, '__call__', recording)
    for n_blocks in (4, 4, 2):
        seen.append([])
        model.blocked_mcmc.sample__(batch_size=3, n_blocks=n_blocks)
    draws = [torch.cat(v).cpu().numpy() for v in seen]
    assert [d.size for d in draws] == [3 * 16, 3 * 16, 3 * 16]
    for i in range(3):
        for j in range(i + 1, 3):                      # no proposal value of one call repeats in another
            assert np.intersect1d(draws[i], draws[j]).size == 0
