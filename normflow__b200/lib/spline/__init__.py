from .spline import RQSpline  # noqa: F401
