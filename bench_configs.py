"""The five workloads of BASELINE.json `configs`, as data.

One builder serves both arms of bench.py: `pkg` is either `normflow__b200` (the CUDA path) or
the staged, unmodified reference (`oracle/_ref/normflow_ref`, CPU float64) -- they expose the
same Python API, so the model of a configuration is assembled by the same calls.
"""

import numpy as np

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)      # examples/scalar_affine.py:14
KNOTS = 10
HIDDEN = [8, 8]
P_OF = {'affine': 2, 'rqs': 3 * KNOTS - 2}

CONFIGS = {
    1: dict(name="configs[0]: 0-dim phi^4, DistConvertor_(10, symmetric)", lattice=(1,),
            blocks=[('distconv', 1)], action=dict(kappa=0, m_sq=-1.2, lambd=0.5), batch=128, cpu_batch=128),
    2: dict(name="configs[1]: 2-D phi^4 16x16, affine coupling x4, ConvAct(1->8->8->2)", lattice=(16, 16),
            blocks=[('affine', 4)], action=ACTION, batch=1024, cpu_batch=1024),
    3: dict(name="configs[2]: 2-D phi^4 64x64, RQ-spline coupling x4 (K=10), ConvAct(1->8->8->28)", lattice=(64, 64),
            blocks=[('rqs', 4)], action=ACTION, batch=16384, cpu_batch=256),
    4: dict(name="configs[3]: 3-D phi^4 32^3, affine x4 + RQ-spline x4", lattice=(32, 32, 32),
            blocks=[('affine', 4), ('rqs', 4)], action=ACTION, batch=4096, batch_is_global=True, cpu_batch=8),
    5: dict(name="configs[4]: 4-D phi^4 16^4, 2 x (affine x4 + RQ-spline x4), Conv4d conditioners",
            lattice=(16, 16, 16, 16), blocks=[('affine', 4), ('rqs', 4)] * 2, action=ACTION, batch=2048, cpu_batch=4),
}


def volume(cfg):
    return int(np.prod(cfg['lattice']))


def fwd_bytes_per_sample(cfg):
    """prior write 4V + per atomic step (8 + 4P)V + action read 4V (the pointwise
    DistConvertor_ of config 1: read + write = 8V)."""
    V = volume(cfg)
    total = 8 * V
    for kind, n in cfg['blocks']:
        total += n * (8 * V if kind == 'distconv' else (8 + 4 * P_OF[kind]) * V)
    return total


def train_bytes_per_sample(cfg):
    """forward + per atomic step (12 + 8P)V for the backward + 8V action backward."""
    V = volume(cfg)
    total = fwd_bytes_per_sample(cfg) + 8 * V
    for kind, n in cfg['blocks']:
        total += n * (16 * V if kind == 'distconv' else (12 + 8 * P_OF[kind]) * V)
    return total


def build_model(pkg, cfg, seed=0):
    """Model of configuration `cfg` out of package `pkg` (normflow__b200 or the reference)."""
    torch = pkg.torch
    nn = pkg.nn
    torch.manual_seed(seed)
    shape = tuple(cfg['lattice'])
    if cfg['blocks'][0][0] == 'distconv':
        net_ = nn.DistConvertor_(KNOTS, symmetric=True)             # examples/scalar_zerodim.py:20
        prior = pkg.prior.NormalPrior(shape=shape[0] if len(shape) == 1 else shape)
    else:
        mask = pkg.mask.EvenOddMask(shape=shape)
        conv = dict(in_channels=1, hidden_sizes=list(HIDDEN), kernel_size=3, conv_dim=len(shape),
                    acts=('tanh', 'tanh', None), bias=False, padding_mode='circular')   # scalar_affine.py:89-99
        blocks = []
        for kind, n in cfg['blocks']:
            nets = [nn.ConvAct(out_channels=P_OF[kind], **conv) for _ in range(n)]
            if kind == 'affine':
                blocks.append(nn.AffineCoupling_(nets, mask=mask))
            else:
                blocks.append(nn.RQSplineCoupling_(nets, mask=mask, xlim=(-5, 5), ylim=(-5, 5),
                                                   extrap=dict(left='linear', right='linear')))
        net_ = nn.ModuleList_(blocks)
        prior = pkg.prior.NormalPrior(shape=shape)
    return pkg.Model(net_=net_, prior=prior, action=pkg.action.ScalarPhi4Action(**cfg['action']))
