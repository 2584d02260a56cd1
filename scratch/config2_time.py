"""Config 2 (16x16, affine x4, B=1024): latency-bound regime."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normflow__b200 import Model, _C
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import ModuleList_, ConvAct, AffineCoupling_
from normflow__b200.prior import NormalPrior
torch.manual_seed(0); np.random.seed(0)
shape, B = (16, 16), 1024
mask = EvenOddMask(shape=shape)
nets = [ConvAct(1, 2, 3, conv_dim=2, hidden_sizes=[8, 8], acts=('tanh', 'tanh', None), bias=False) for _ in range(4)]
model = Model(prior=NormalPrior(shape=shape), net_=ModuleList_([AffineCoupling_(nets, mask=mask)]),
              action=ScalarPhi4Action(kappa=0.67, m_sq=-2.68, lambd=0.5))
model.device_handler.to('cuda')
model.fit(n_epochs=5, batch_size=B, checkpoint_dict=dict(print_stride=1000, display=False))
torch.cuda.synchronize()
for name, fn, n in [("fit.step", model.fit.step, 200), ("posterior.sample__", lambda: model.posterior.sample__(B), 500),
                    ("mcmc.sample", lambda: model.mcmc.sample(B), 200)]:
    fn(); torch.cuda.synchronize()
    n0 = _C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:20s} wall {wall:.3f} ms/call  gpu-span {e0.elapsed_time(e1) / n:.3f} ms  -> {B / wall * 1e3:.0f} samples/s; {( _C.launch_count() - n0) / n:.1f} nfk launches/call")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(300): model.posterior.sample__(B)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
