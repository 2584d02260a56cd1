from .prior import Prior, NormalPrior  # noqa: F401
