"""Containers for invertible layers (reference src/nn/_core.py).

A trailing underscore marks layers whose `forward(x, log0)` and `backward(x, log0)`
return the transformed field together with the accumulated log-Jacobian.
"""

import base64
import copy
import io

import numpy as np
import torch


class Module_(torch.nn.Module):
    """Base class of an invertible layer (reference nn/_core.py:12-42)."""

    propagate_density = False

    def __init__(self, label=None):
        super().__init__()
        self.label = label

    def forward(self, x, log0=0):
        raise NotImplementedError

    def backward(self, x, log0=0):
        raise NotImplementedError

    def transfer(self, **kwargs):
        return copy.deepcopy(self)

    @property
    def npar(self):
        return sum(int(np.prod(p.shape)) for p in self.parameters())

    def sum_density(self, x):
        """Sum a per-site log-Jacobian over everything but the batch axis."""
        if self.propagate_density:
            return x
        return torch.sum(x, dim=list(range(1, x.dim())))


class ModuleList_(torch.nn.ModuleList):
    """A chain of invertible layers (reference nn/_core.py:46-134): `forward` applies
    them in order, `backward` applies their inverses in reverse order."""

    _groups = None

    def __init__(self, nets_, label=None):
        super().__init__(nets_)
        self.label = label

    def forward(self, x, log0=0):
        for net_ in self:
            x, log0 = net_.forward(x, log0)
        return x, log0

    def backward(self, x, log0=0):
        for net_ in reversed(list(self)):
            x, log0 = net_.backward(x, log0)
        return x, log0

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def hack(self, x, log0=0):
        """forward() that also returns every intermediate (x, log) pair."""
        stack = [(x, log0)]
        for net_ in self:
            x, log0 = net_.forward(x, log0)
            stack.append((x, log0))
        return stack

    # --- parameter groups for the optimiser (nn/_core.py:77-93) ---------------------
    def setup_groups(self, groups=None):
        """groups = [{'ind': [0, 1], 'hyper': dict(weight_decay=1e-4)}, ...]"""
        self._groups = groups

    def grouped_parameters(self):
        if self._groups is None:
            return super().parameters()
        out = []
        for grp in self._groups:
            params = [p for k in grp['ind'] for p in self[k].parameters()]
            out.append(dict(params=params, **grp['hyper']))
        return out

    def transfer(self, **kwargs):
        return self.__class__([net_.transfer(**kwargs) for net_ in self])

    # --- weights as a text blob (nn/_core.py:108-118) ------------------------------
    def get_weights_blob(self):
        buf = io.BytesIO()
        torch.save(self.state_dict(), buf)
        return base64.b64encode(buf.getbuffer()).decode('utf-8')

    def set_weights_blob(self, blob, map_location=torch.device('cpu')):
        state = torch.load(io.BytesIO(base64.b64decode(blob.strip())), map_location=map_location)
        self.load_state_dict(state)

    def freeze_parameters(self):
        for p in self.parameters():
            p.requires_grad = False

    def unfreeze_parameters(self):
        for p in self.parameters():
            p.requires_grad = True

    @property
    def npar(self):
        return sum(int(np.prod(p.shape)) for p in super().parameters())

    def to(self, *args, **kwargs):
        for net_ in self:
            net_.to(*args, **kwargs)
        return self
