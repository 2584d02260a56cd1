"""Site masks that split a lattice field into two partitions (reference src/mask/mask.py).

A mask offers `split`, `cat` and `purify`.  The coupling layers of this package do not
call them on their fast path (the partition test is fused into the coupling and the
first conditioner layer); they stay available, kernel-backed, for user code and for
coupling subclasses that follow the reference's split -> atomic -> cat dataflow.
"""

import torch

from .. import _ops


class Mask(torch.nn.Module):
    """0/1 site mask held as uint8 buffers `_mask` and `_c_mask` (same state_dict
    entries as the reference, mask.py:17-24)."""

    def __init__(self, **mask_kwargs):
        super().__init__()
        mask = self.make_mask(**mask_kwargs)
        self.register_buffer('_mask', mask)
        self.register_buffer('_c_mask', 1 - mask)
        self.mask_kwargs = mask_kwargs

    def __str__(self):
        return str(self._mask)

    @property
    def shape(self):
        return tuple(self._mask.shape)

    def split(self, x):
        """(partition 0, partition 1), each zero outside its sites (mask.py:30-31)."""
        return _ops.mask_select(x, self._mask, 1), _ops.mask_select(x, self._mask, 0)

    def cat(self, x_0, x_1):
        return x_0 + x_1

    def purify(self, x_chnl, channel):
        """Zero everything outside partition `channel` (mask.py:36-37)."""
        return _ops.mask_select(x_chnl, self._mask, 1 if channel == 0 else 0)

    @staticmethod
    def make_mask():
        raise NotImplementedError


def _coordinate_sum(shape, skip=None, only=None):
    """Sum of site coordinates as an int64 tensor of the lattice shape."""
    total = torch.zeros(tuple(shape), dtype=torch.int64, device='cpu')
    for axis, extent in enumerate(shape):
        if axis == skip or (only is not None and axis != only):
            continue
        view = [1] * len(shape)
        view[axis] = extent
        total = total + torch.arange(extent, dtype=torch.int64, device='cpu').view(view)
    return total


class EvenOddMask(Mask):
    """Checkerboard: (1 - parity + sum of coordinates [minus the `exclude_mu` one]) mod 2
    (reference mask.py:53-61; bit-exact)."""

    @staticmethod
    def make_mask(*, shape, parity=0, exclude_mu=None):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        if exclude_mu is not None and exclude_mu < 0:
            exclude_mu += len(shape)
        bits = torch.remainder(1 - parity + _coordinate_sum(shape, skip=exclude_mu), 2)
        return bits.to(torch.uint8).to(torch.get_default_device())


class AlongAxesEvenOddMask(Mask):
    """Alternates along one axis only (reference mask.py:64-72)."""

    @staticmethod
    def make_mask(*, shape, parity=0, mu=0):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        if mu < 0:
            mu += len(shape)
        bits = torch.remainder(1 - parity + _coordinate_sum(shape, only=mu), 2)
        return bits.to(torch.uint8).to(torch.get_default_device())


class DummyMask:
    """Everything in one partition (reference mask.py:75-94)."""

    def __init__(self, parity=0):
        self.parity = parity

    def split(self, x):
        return (x, None) if self.parity == 0 else (None, x)

    def cat(self, x_0, x_1):
        return x_0 if self.parity == 0 else x_1

    @staticmethod
    def purify(x_chnl, *args, **kwargs):
        return x_chnl
