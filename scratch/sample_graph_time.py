"""posterior.sample__ in the launch-bound configurations: eager vs CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import normflow__b200 as nf
from normflow__b200 import Model
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import ModuleList_, ConvAct, AffineCoupling_, DistConvertor_
from normflow__b200.prior import NormalPrior

def time_it(model, B, n=300):
    for _ in range(10):
        model.posterior.sample__(B)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n):
        model.posterior.sample__(B)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n

torch.manual_seed(0)
shape = (16, 16)
conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, conv_dim=2, acts=('tanh', 'tanh', None), bias=False)
cfg2 = Model(prior=NormalPrior(shape=shape), net_=ModuleList_([AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=shape))]),
             action=ScalarPhi4Action(kappa=0.67, m_sq=-2.68, lambd=0.5))
cfg1 = Model(prior=NormalPrior(shape=1), net_=DistConvertor_(10, symmetric=True), action=ScalarPhi4Action(kappa=0, m_sq=-1.2, lambd=0.5))
for name, model, B in (("config 2 (16x16 affine x4)", cfg2, 1024), ("config 1 (0-dim)", cfg1, 128), ("config 1 (0-dim)", cfg1, 16384)):
    model.device_handler.to('cuda')
    res = {}
    for g in (False, True):
        model.posterior.cuda_graph = g
        res[g] = time_it(model, B)
    print(f"{name} B={B}: eager {res[False] * 1e6:.0f} us -> {B / res[False]:.3g} samples/s; graph {res[True] * 1e6:.0f} us -> {B / res[True]:.3g} samples/s")
