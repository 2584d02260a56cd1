"""examples/scalar_affine.py's net on one B200: training epochs (eager / CUDA graph), posterior
sampling, blocked MCMC.  Counterpart of scratch/ref_example_time.py (reference on CPU)."""
import json, sys, time
import numpy as np
import torch
import normflow__b200 as nf
from normflow__b200 import Model, _C
from normflow__b200.nn import *
from normflow__b200.mask import EvenOddMask
from normflow__b200.prior import NormalPrior
from normflow__b200.action import ScalarPhi4Action

out = {}
for lat, B in [((8, 8), 128), ((8, 8), 4096), ((64, 64), 4096)]:
    res = {}
    for graph in (False, True):
        torch.manual_seed(0); np.random.seed(0)
        mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
        ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
        conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, padding_mode='circular',
                    conv_dim=len(lat), acts=('tanh', 'tanh', None), bias=False)
        net_ = ModuleList_([PSDBlock_(mfnet_=mf, fftnet_=ff), DistConvertor_(50, symmetric=True, smooth=True),
                            AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
                            DistConvertor_(50, symmetric=True, smooth=True)])
        model = Model(net_=net_, prior=NormalPrior(shape=lat),
                      action=ScalarPhi4Action(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5))
        model.device_handler.to("cuda")
        model.fit.cuda_graph = graph
        model.fit(n_epochs=20, batch_size=B, checkpoint_dict=dict(print_stride=1000))     # warm-up
        torch.cuda.synchronize()
        n = 200
        t = time.time()
        model.fit(n_epochs=n, batch_size=B, checkpoint_dict=dict(print_stride=1000))
        torch.cuda.synchronize()
        dt = time.time() - t
        res["train_graph_samples_per_s" if graph else "train_eager_samples_per_s"] = n * B / dt
        res["train_graph_ms_per_epoch" if graph else "train_eager_ms_per_epoch"] = dt / n * 1e3
    with torch.no_grad():
        S = max(B, 1024)
        model.posterior.sample__(S)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(20):
            model.posterior.sample__(S)
        torch.cuda.synchronize()
        res["posterior_samples_per_s"] = 20 * S / (time.time() - t)
    if lat == (8, 8) and B == 128:
        model.blocked_mcmc.sample__(batch_size=4, n_blocks=4)
        torch.cuda.synchronize()
        n0 = _C.launch_count()
        t = time.time()
        model.blocked_mcmc.sample__(batch_size=64, n_blocks=4)
        torch.cuda.synchronize()
        dt = time.time() - t
        res["blocked_mcmc_block_updates_per_s"] = 64 * 4 / dt
        res["blocked_mcmc_launches_per_update"] = (_C.launch_count() - n0) / (64 * 4)
    out[f"{'x'.join(map(str, lat))}_B{B}"] = res
    print(lat, B, json.dumps(res), flush=True)
json.dump(out, open(sys.argv[1], "w"), indent=1)
