import sys
import numpy as np, torch
import normflow__b200 as nf
from normflow__b200 import Model
from normflow__b200.nn import *
from normflow__b200.mask import EvenOddMask
from normflow__b200.prior import NormalPrior
from normflow__b200.action import ScalarPhi4Action
lat = tuple(int(v) for v in sys.argv[1].split(",")); B = int(sys.argv[2]); n = int(sys.argv[3])
torch.manual_seed(0); np.random.seed(0)
mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, padding_mode='circular',
            conv_dim=len(lat), acts=('tanh', 'tanh', None), bias=False)
net_ = ModuleList_([PSDBlock_(mfnet_=mf, fftnet_=ff), DistConvertor_(50, symmetric=True, smooth=True),
                    AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
                    DistConvertor_(50, symmetric=True, smooth=True)])
model = Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5))
model.device_handler.to("cuda")
model.fit(n_epochs=n, batch_size=B, checkpoint_dict=dict(print_stride=1000))
torch.cuda.synchronize()
