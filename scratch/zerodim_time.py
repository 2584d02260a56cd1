"""The reference's only published timing: 0-dim phi^4, DistConvertor_(10, symmetric=True), 1000 epochs x batch 1024
(examples/0dim_normflow_example.ipynb: 2.91 s on a CPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from normflow__b200 import Model, _C
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.prior import NormalPrior
from normflow__b200.nn import DistConvertor_
torch.manual_seed(0); np.random.seed(0)
model = Model(prior=NormalPrior(shape=(1,)), net_=DistConvertor_(10, symmetric=True),
              action=ScalarPhi4Action(kappa=0, m_sq=-1.2, lambd=0.5))
model.device_handler.to('cuda')
model.fit(n_epochs=20, batch_size=1024, checkpoint_dict=dict(print_stride=1000, display=False))
torch.cuda.synchronize()
n0 = _C.launch_count()
t0 = time.time()
model.fit(n_epochs=1000, batch_size=1024, checkpoint_dict=dict(print_stride=100))
torch.cuda.synchronize()
dt = time.time() - t0
print(f"1000 epochs x 1024: {dt:.2f} s  ({1024 * 1000 / dt:.0f} train samples/s), {(_C.launch_count() - n0) / 1000:.1f} nfk launches/epoch")
t0 = time.time()
for _ in range(200):
    model.fit.step()
torch.cuda.synchronize()
print(f"bare fit.step(): {(time.time() - t0) / 200 * 1e3:.3f} ms")
