"""Configs 4 and 5 at full size: do the N-D (unfused) kernels hold up (indexing beyond 2^31 elements, memory)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normflow__b200 import Model
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import ModuleList_, ConvAct, AffineCoupling_, RQSplineCoupling_
from normflow__b200.prior import NormalPrior
def build(shape, n_aff, n_rqs, repeat=1):
    torch.manual_seed(0)
    mask = EvenOddMask(shape=shape)
    conv = dict(in_channels=1, hidden_sizes=[8, 8], kernel_size=3, conv_dim=len(shape), acts=('tanh', 'tanh', None), bias=False)
    blocks = []
    for _ in range(repeat):
        blocks.append(AffineCoupling_([ConvAct(out_channels=2, **conv) for _ in range(n_aff)], mask=mask))
        blocks.append(RQSplineCoupling_([ConvAct(out_channels=28, **conv) for _ in range(n_rqs)], mask=mask, xlim=(-5, 5), ylim=(-5, 5),
                                        extrap=dict(left='linear', right='linear')))
    m = Model(prior=NormalPrior(shape=shape), net_=ModuleList_(blocks), action=ScalarPhi4Action(kappa=0.67, m_sq=-2.68, lambd=0.5))
    m.device_handler.to('cuda')
    return m
for name, shape, B, args in [("config 4: 32^3, affine x4 + RQS x4", (32, 32, 32), 4096, (4, 4, 1)),
                             ("config 5: 16^4, 2 x (affine x4 + RQS x4), Conv4d", (16, 16, 16, 16), 2048, (4, 4, 2))]:
    model = build(shape, *args)
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    for rep in range(2):
        t0 = time.time()
        y, logq, logp = model.posterior.sample__(B)
        torch.cuda.synchronize(); dt = time.time() - t0
    ok = bool(torch.isfinite(y).all() and torch.isfinite(logq).all() and torch.isfinite(logp).all())
    # inverse round trip on a slice
    with torch.no_grad():
        xb, lb = model.net_.backward(y[:64], log0=torch.zeros(64, device='cuda'))
        y2, l2 = model.net_(xb)
    rt = float((y2 - y[:64]).abs().max())
    print(f"### {name}: sample__({B}) {dt * 1e3:.0f} ms -> {B / dt:.0f} samples/s; finite {ok}; fwd(inv(y)) - y max {rt:.2e}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del model, y, logq, logp; torch.cuda.empty_cache()
