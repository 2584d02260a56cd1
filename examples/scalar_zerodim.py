"""Zero-dimensional phi^4 "lattice" (one site): train a DistConvertor_ flow towards
exp(-S), S = m_sq/2 phi^2 + lambd phi^4 -- the workload of the reference's
examples/scalar_zerodim.py, on this package.

    python examples/scalar_zerodim.py --n_epochs 1000 --batch_size 1024 [--graph]

The loss converges to -log Z = -1.112773 for the default couplings.
"""
import argparse

from normflow__b200 import Model, backward_sanitychecker
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.nn import DistConvertor_
from normflow__b200.prior import NormalPrior


def train(model, **fit_kwargs):
    model.fit(**fit_kwargs)


def main(m_sq=-1.2, lambd=0.5, knots_len=10, n_epochs=1000, batch_size=1024, lat_shape=1, nranks=1,
         graph=False, snapshot_path=None, print_stride=100):
    model = Model(net_=DistConvertor_(knots_len, symmetric=True),
                  prior=NormalPrior(shape=lat_shape),
                  action=ScalarPhi4Action(kappa=0, m_sq=m_sq, lambd=lambd))
    print("number of model parameters =", model.net_.npar)
    model.fit.cuda_graph = graph and nranks == 1
    fit_kwargs = dict(n_epochs=n_epochs, save_every=None, batch_size=batch_size // nranks,
                      hyperparam=dict(lr=0.01, weight_decay=0., fused=True),
                      checkpoint_dict=dict(print_stride=print_stride, snapshot_path=snapshot_path))
    if nranks > 1:
        model.device_handler.spawnprocesses(train, nranks, **fit_kwargs)
    else:
        model.fit(**fit_kwargs)
    backward_sanitychecker(model)
    return model


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument("--m_sq", type=float, default=-1.2)
    ap.add_argument("--lambd", type=float, default=0.5)
    ap.add_argument("--knots_len", type=int, default=10)
    ap.add_argument("--n_epochs", type=int, default=1000)
    ap.add_argument("--batch_size", type=int, default=1024)
    ap.add_argument("--nranks", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="CUDA-graph training loop")
    main(**vars(ap.parse_args()))
