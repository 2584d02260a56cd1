from .resampler import Resampler  # noqa: F401
