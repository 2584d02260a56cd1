"""phi^4 on a 2-D (or 3-D / 4-D) lattice with the net of the reference's examples/scalar_affine.py:
PSDBlock_ (mean-field net + FFTNet_) -> DistConvertor_ -> AffineCoupling_ x n_layers (ConvAct
conditioners on a checkerboard mask) -> DistConvertor_, trained on the reverse KL divergence.

    python examples/scalar_affine.py --lat_shape 8,8 --n_epochs 1000 --batch_size 128 [--graph]
"""
import argparse

from normflow__b200 import Model, backward_sanitychecker
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import (ModuleList_, Identity_, DistConvertor_, AffineCoupling_, ConvAct,
                               FFTNet_, MeanFieldNet_, PSDBlock_)
from normflow__b200.prior import NormalPrior


def assemble_net(*, lat_shape, n_layers=4, hidden_sizes=(8, 8), zee2sym=True, acts=None,
                 knots0_len=10, knots1_len=10, knots2_len=50, knots4_len=50):
    """zee2sym: keep the Z2 symmetry phi -> -phi (odd activations, no biases, symmetric splines)."""
    blocks = []
    mfnet_ = (MeanFieldNet_.build(knots_len=knots0_len, symmetric=zee2sym, final_scale=True, smooth=True)
              if knots0_len > 1 else Identity_())
    fftnet_ = FFTNet_.build(lat_shape, knots_len=knots1_len, ignore_zeromode=True)
    blocks.append(PSDBlock_(mfnet_=mfnet_, fftnet_=fftnet_))
    if knots2_len > 1:
        blocks.append(DistConvertor_(knots2_len, symmetric=zee2sym, smooth=True))
    if acts is None:
        acts = (*(['tanh' if zee2sym else 'leaky_relu'] * len(hidden_sizes)), None)
    conv = dict(in_channels=1, out_channels=2, hidden_sizes=list(hidden_sizes), kernel_size=3,
                padding_mode='circular', conv_dim=len(lat_shape), acts=acts, bias=not zee2sym)
    blocks.append(AffineCoupling_([ConvAct(**conv) for _ in range(n_layers)], mask=EvenOddMask(shape=lat_shape)))
    if knots4_len > 1:
        blocks.append(DistConvertor_(knots4_len, symmetric=zee2sym, smooth=True))
    return ModuleList_(blocks)


def train(model, **fit_kwargs):
    model.fit(**fit_kwargs)


def main(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5, n_epochs=1000, batch_size=128, lat_shape=(8, 8), nranks=1,
         graph=False, snapshot_path=None, print_stride=100, save_every=200, **net_kwargs):
    model = Model(net_=assemble_net(lat_shape=lat_shape, **net_kwargs),
                  prior=NormalPrior(shape=lat_shape),
                  action=ScalarPhi4Action(kappa=kappa, m_sq=m_sq, lambd=lambd))
    print("number of model parameters =", model.net_.npar)
    # weight decay per block: light on the spectral / pointwise blocks, heavier on the couplings
    model.net_.setup_groups(groups=[{'ind': [0, 1, 3], 'hyper': dict(weight_decay=1e-4)},
                                    {'ind': [2], 'hyper': dict(weight_decay=1e-2)}])
    model.fit.cuda_graph = graph and nranks == 1
    fit_kwargs = dict(n_epochs=n_epochs, save_every=save_every, batch_size=batch_size // nranks,
                      hyperparam=dict(fused=True),
                      checkpoint_dict=dict(print_stride=print_stride, snapshot_path=snapshot_path))
    if nranks > 1:
        model.device_handler.spawnprocesses(train, nranks, **fit_kwargs)
    else:
        model.fit(**fit_kwargs)
    backward_sanitychecker(model)
    return model


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument("--lat_shape", type=lambda s: tuple(int(v) for v in s.split(",")), default=(8, 8))
    ap.add_argument("--m_sq", type=float, default=-4 * 0.67)
    ap.add_argument("--lambd", type=float, default=0.5)
    ap.add_argument("--kappa", type=float, default=0.67)
    ap.add_argument("--n_epochs", type=int, default=1000)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--nranks", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="CUDA-graph training loop")
    main(**vars(ap.parse_args()))
