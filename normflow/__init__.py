"""`normflow` -- the reference's package name, served by normflow__b200.

    from normflow import Model, np, torch, backward_sanitychecker
    from normflow.nn import DistConvertor_
    from normflow.action import ScalarPhi4Action

Existing scripts written against jkomijani/normflow_ (examples/scalar_zerodim.py:1-5,
examples/scalar_affine.py:2-9, src/__init__.py:4-13) import this name; every submodule of
`normflow` IS the module of the same name in `normflow__b200` (one set of classes, not a copy).
"""

import importlib
import sys

import normflow__b200 as _impl
from normflow__b200 import *  # noqa: F401,F403
from normflow__b200 import Model, np, torch, backward_sanitychecker  # noqa: F401

__all__ = list(_impl.__all__)

# `normflow.nn`, `normflow.nn.scalar.couplings_`, ... resolve to the normflow__b200 modules themselves
for _name, _mod in list(sys.modules.items()):
    if _name == "normflow__b200" or _name.startswith("normflow__b200."):
        sys.modules.setdefault("normflow" + _name[len("normflow__b200"):], _mod)


def __getattr__(name):
    """Submodules not imported yet (`normflow._build`, ...)."""
    try:
        mod = importlib.import_module("normflow__b200." + name)
    except ImportError as err:
        raise AttributeError(f"module 'normflow' has no attribute '{name}'") from err
    sys.modules.setdefault("normflow." + name, mod)
    return mod
