"""Time the REFERENCE on this container's CPU cores (scratch measurement, needs /root/reference):
examples/scalar_affine.py's net on 8x8 -- training epochs, posterior sampling, blocked MCMC."""
import os, sys, tempfile, time, warnings
import numpy as np
warnings.filterwarnings("ignore"); np.product = np.prod
_tmp = tempfile.mkdtemp(); os.symlink("/root/reference/src", os.path.join(_tmp, "normflow_ref")); sys.path.insert(0, _tmp)
import torch
import normflow_ref as nf
from normflow_ref import Model
from normflow_ref.nn import *
from normflow_ref.mask import EvenOddMask
from normflow_ref.prior import NormalPrior
from normflow_ref.action import ScalarPhi4Action

lat = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8,8").split(","))
mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, padding_mode='circular',
            conv_dim=len(lat), acts=('tanh', 'tanh', None), bias=False)
net_ = ModuleList_([PSDBlock_(mfnet_=mf, fftnet_=ff), DistConvertor_(50, symmetric=True, smooth=True),
                    AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
                    DistConvertor_(50, symmetric=True, smooth=True)])
model = Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5))
print("threads", torch.get_num_threads())
t = time.time(); model.fit(n_epochs=100, batch_size=128, checkpoint_dict=dict(print_stride=100)); dt = time.time() - t
print(f"train: 100 epochs x 128 in {dt:.2f} s -> {100 * 128 / dt:.0f} samples/s")
with torch.no_grad():
    model.posterior.sample__(1024)
    t = time.time()
    for _ in range(5): model.posterior.sample__(1024)
    dt = time.time() - t
print(f"posterior.sample__: {5 * 1024 / dt:.0f} samples/s")
t = time.time(); model.blocked_mcmc.sample__(batch_size=16, n_blocks=4); dt = time.time() - t
print(f"blocked_mcmc: 16 samples x 4 blocks in {dt:.2f} s -> {16 * 4 / dt:.1f} block updates/s")
