from . import combo, stats, spline  # noqa: F401
