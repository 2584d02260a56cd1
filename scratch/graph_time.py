import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normflow__b200 import Model
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.prior import NormalPrior
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import DistConvertor_, ModuleList_, ConvAct, AffineCoupling_
def zero_dim():
    return Model(prior=NormalPrior(shape=(1,)), net_=DistConvertor_(10, symmetric=True), action=ScalarPhi4Action(kappa=0, m_sq=-1.2, lambd=0.5))
def config2():
    mask = EvenOddMask(shape=(16, 16))
    nets = [ConvAct(1, 2, 3, conv_dim=2, hidden_sizes=[8, 8], acts=('tanh', 'tanh', None), bias=False) for _ in range(4)]
    return Model(prior=NormalPrior(shape=(16, 16)), net_=ModuleList_([AffineCoupling_(nets, mask=mask)]), action=ScalarPhi4Action(kappa=0.67, m_sq=-2.68, lambd=0.5))
for name, make in [("0-dim (reference notebook: 2.91 s on CPU)", zero_dim), ("config 2: 16x16 affine x4", config2)]:
    for graph in (False, True):
        torch.manual_seed(0); np.random.seed(0)
        model = make(); model.device_handler.to('cuda'); model.fit.cuda_graph = graph
        model.fit(n_epochs=10, batch_size=1024, checkpoint_dict=dict(print_stride=10**6))      # warm everything
        torch.cuda.synchronize(); t0 = time.time()
        model.fit(n_epochs=1000, batch_size=1024, checkpoint_dict=dict(print_stride=100))
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"### {name}: graph={graph}: 1000 epochs x 1024 in {dt:.2f} s ({1024e3 / dt:.0f} train samples/s), final loss {np.mean(model.fit.train_history['loss'][-20:]):.4f}")
