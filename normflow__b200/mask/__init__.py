from .mask import Mask, EvenOddMask, AlongAxesEvenOddMask, DummyMask  # noqa: F401
