"""normflow__b200 -- B200-native implementation of the data-parallel hot path of
jkomijani/normflow_ behind the reference's own Python API.

    from normflow__b200 import Model, np, torch, backward_sanitychecker
    from normflow__b200 import action, mask, nn, prior, mcmc

mirrors `normflow/__init__.py` (reference src/__init__.py:4-13).  The inner loop of
Model.fit / posterior.sample / mcmc.sample runs in hand-written sm_100a CUDA kernels
(libnormflow_b200.so, C ABI in include/normflow_b200.h); PyTorch supplies device
memory, streams, autograd bookkeeping and torch.distributed.  There is no CPU path.
"""

from . import device  # noqa: F401  (sets the default device like the reference does)

from ._normflowcore import Model
from ._normflowcore import np, torch
from ._normflowcore import backward_sanitychecker

from . import action
from . import mask
from . import nn
from . import prior
from . import mcmc
from . import lib

__all__ = ["Model", "np", "torch", "backward_sanitychecker",
           "action", "mask", "nn", "prior", "mcmc", "lib", "device"]
