"""Controlled couplings (reference src/nn/scalar/cntr_couplings_.py): the first atomic step is
conditioned on an external `control` field instead of the frozen partition of the data."""

from ... import _C
from .couplings_ import Coupling_, ShiftCoupling_, AffineCoupling_, RQSplineCoupling_, MultiRQSplineCoupling_


class DirectCntrCoupling_(Coupling_):
    """forward((x, control), log0) -> ((y, control), logJ)  (cntr_couplings_.py:17-52).

    Step 0 feeds `control` (the whole field, unmasked) to its conditioner; the other steps see
    the frozen partition of the data as usual.  Couplings with a full-field kernel
    (`_transform`) run it directly; other subclasses go through split / atomic_* / cat."""

    def _full_field(self):
        return type(self)._transform is not Coupling_._transform

    def _step_out(self, k, net, x, control, parity):
        if k == 0:
            return net(self.preprocess_fz(control))
        if isinstance(self, MultiRQSplineCoupling_):
            return net(self.preprocess_fz(self._frozen(x, parity)))
        return self._conditioner(net, x, parity)

    def _controlled(self, x_and_control, log0, inverse):
        x, control = x_and_control
        order = range(len(self.nets))
        if self._full_field():
            for k in (reversed(order) if inverse else order):
                p = k % 2
                out = self._step_out(k, self.nets[k], x, control, p)
                x, log0 = self._transform(x, out, p, log0, _C.FROZEN_COPY, inverse)
            return (x, control), log0
        parts = list(self.mask.split(x))
        for k in (reversed(order) if inverse else order):
            p = k % 2
            step = self.atomic_backward if inverse else self.atomic_forward
            parts[p], log0 = step(x_active=parts[p], x_frozen=control if k == 0 else parts[1 - p],
                                  parity=p, net=self.nets[k], log0=log0)
        return (self.mask.cat(*parts), control), log0

    def forward(self, x_and_control, log0=0):
        return self._controlled(x_and_control, log0, inverse=False)

    def backward(self, x_and_control, log0=0):
        return self._controlled(x_and_control, log0, inverse=True)


class CntrCoupling_(DirectCntrCoupling_):
    """The control field comes from `control_generator(batch_size)` at every forward call and is
    kept (as `.control`) for the matching backward call; the caller only sees the data
    (cntr_couplings_.py:56-83)."""

    def __init__(self, *args, control_generator=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.control_generator = control_generator
        self.control = None

    def forward(self, x, log0=0):
        self.control = self.control_generator(x.shape[0])
        (x, _), log0 = DirectCntrCoupling_.forward(self, (x, self.control), log0=log0)
        return x, log0

    def backward(self, x, log0=0):
        (x, _), log0 = DirectCntrCoupling_.backward(self, (x, self.control), log0=log0)
        return x, log0


class CntrShiftCoupling_(CntrCoupling_, ShiftCoupling_):
    pass


class CntrAffineCoupling_(CntrCoupling_, AffineCoupling_):
    pass


class CntrRQSplineCoupling_(CntrCoupling_, RQSplineCoupling_):
    pass


class CntrMultiRQSplineCoupling_(CntrCoupling_, MultiRQSplineCoupling_):
    pass
