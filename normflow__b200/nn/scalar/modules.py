"""Plain (Jacobian-free) networks: the ConvAct conditioner and the SplineNet knot
parameterisation (reference src/nn/scalar/modules.py, convNd.py).

The convolution layers subclass torch's Conv modules only to inherit their parameter
names, shapes and initialisation (so state_dicts are interchangeable with the
reference); the arithmetic is the package's own circular-convolution kernel, which
wraps indices instead of materialising a padded copy.
"""

import copy

import torch

from ... import _ops
from ...lib.spline import RQSpline


def _check_conv(conv):
    ks = conv.kernel_size
    if len(set(ks)) != 1 or ks[0] % 2 != 1:
        raise NotImplementedError("circular conv kernel: one odd kernel size in every direction")
    if conv.padding != 'same' and tuple(conv.padding) != tuple(k // 2 for k in ks):
        raise NotImplementedError("circular conv kernel: padding must be 'same'")
    if conv.padding_mode != 'circular':
        raise NotImplementedError("circular conv kernel: padding_mode must be 'circular'")
    if tuple(conv.stride) != (1,) * len(ks) or tuple(conv.dilation) != (1,) * len(ks) or conv.groups != 1:
        raise NotImplementedError("circular conv kernel: stride = dilation = groups = 1")


class _CircularConvMixin:
    """forward() through the package kernel; `standard_weight()` = (Co, Ci, *k)."""

    def standard_weight(self):
        return self.weight

    def forward(self, inp):
        return _ops.conv_stack(inp, [self.standard_weight()], [self.bias], [None], self.kernel_size[0])


class CircularConv1d(_CircularConvMixin, torch.nn.Conv1d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        _check_conv(self)


class CircularConv2d(_CircularConvMixin, torch.nn.Conv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        _check_conv(self)


class CircularConv3d(_CircularConvMixin, torch.nn.Conv3d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        _check_conv(self)


class Conv4d(torch.nn.Module):
    """4-D circular convolution (reference convNd.py:7-149).

    The reference builds it from a Conv3d over (B*L0) with Co*k0 output channels and
    k0 rolled partial sums; its parameter therefore lives in `_conv_lower_dim.weight`
    with shape (Co*k0, Ci, k, k, k), and a bias initialised with randn.  The same
    parameter layout is kept here (state_dict compatible); the computation is one direct
    81-tap kernel on the standard (Co, Ci, k0, k, k, k) view of that weight.
    """

    def __init__(self, in_channels, out_channels, kernel_size, *, stride=1, padding='same',
                 padding_mode='circular', dilation=1, groups=1, bias=True, device=None, dtype=None):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = [kernel_size] * 4
        kernel_size = list(kernel_size)
        if stride != 1 or dilation != 1 or groups != 1 or padding_mode != 'circular':
            raise NotImplementedError("Conv4d: stride = dilation = groups = 1, circular padding only")
        if isinstance(padding, str):
            if padding != 'same':
                raise NotImplementedError("Conv4d: padding must be 'same'")
            lower_padding = padding
        else:
            padding = (padding,) * 4 if isinstance(padding, int) else tuple(padding)
            if padding[0] != kernel_size[0] // 2:
                raise NotImplementedError("Conv4d: padding must be 'same'")
            lower_padding = padding[1:]
        self.conv_ndim = 4
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        # parameter container only: never called
        self._conv_lower_dim = torch.nn.Conv3d(in_channels, out_channels * kernel_size[0], kernel_size[1:],
                                               padding=lower_padding, padding_mode=padding_mode, bias=False,
                                               device=device, dtype=dtype)
        if len(set(kernel_size)) != 1 or kernel_size[0] % 2 != 1:
            raise NotImplementedError("Conv4d: one odd kernel size in every direction")
        self.bias = torch.nn.Parameter(torch.randn(out_channels, dtype=dtype, device=device)) if bias else None

    @property
    def weight(self):
        """(Co, Ci, k0, k, k, k) view of the stored weight (convNd.py:129-142)."""
        k = self.kernel_size
        w = self._conv_lower_dim.weight.reshape(self.out_channels, k[0], self.in_channels, *k[1:])
        return w.movedim(1, 2)

    def standard_weight(self):
        return self.weight.contiguous()

    def forward(self, inp):
        if inp.dim() == 5:
            inp = inp.unsqueeze(0)
        return _ops.conv_stack(inp, [self.standard_weight()], [self.bias], [None], self.kernel_size[0])


class Abs(torch.nn.Module):
    def forward(self, x):
        return torch.abs(x)


class AvgNeighborPool(torch.nn.Module):
    """Mean of the 2 D nearest neighbours (reference lib/linalg/mean.py:7-19)."""

    def forward(self, x):
        dims = range(1, x.ndim)
        return sum(torch.roll(x, s, d) for d in dims for s in (1, -1)) / (2 * len(dims))


ACTIVATIONS = torch.nn.ModuleDict([
    ['tanh', torch.nn.Tanh()],
    ['relu', torch.nn.ReLU()],
    ['leaky_relu', torch.nn.LeakyReLU()],
    ['softplus', torch.nn.Softplus()],
    ['avg_neighbor_pool', AvgNeighborPool()],
    ['abs', Abs()],
    ['expit', torch.nn.Sigmoid()],
    ['none', torch.nn.Identity()],
])

_FUSED_ACTS = (None, 'none', 'tanh', 'relu', 'leaky_relu', 'softplus')


class ConvAct(torch.nn.Sequential):
    """Stack of circular 'same' convolutions with activations: the conditioner
    (reference modules.py:68-154).

        ConvAct(in_channels, out_channels, kernel_size, conv_dim=2, hidden_sizes=[8, 8],
                acts=['tanh', 'tanh', None], bias=False)

    maps (B, in_channels, *L) to (B, out_channels, *L).  `len(acts)` must be
    `len(hidden_sizes) + 1`.  Module indices (hence state_dict keys `0.weight`,
    `2.weight`, ...) follow the reference: every conv is followed by its activation
    module when that is not None.
    """

    Conv = {1: CircularConv1d, 2: CircularConv2d, 3: CircularConv3d, 4: Conv4d}

    def __init__(self, in_channels, out_channels, kernel_size, conv_dim=2, hidden_sizes=[],
                 acts=[None], pre_act=None, **extra_kwargs):
        Conv = self.Conv[conv_dim]
        sizes = [in_channels, *hidden_sizes, out_channels]
        if len(acts) != len(hidden_sizes) + 1:
            raise AssertionError("need one activation entry per conv layer")
        conv_kwargs = dict(padding='same', padding_mode='circular')
        conv_kwargs.update(extra_kwargs)

        layers = [] if pre_act is None else [ACTIVATIONS[pre_act]]
        for i, act in enumerate(acts):
            layers.append(Conv(sizes[i], sizes[i + 1], kernel_size, **conv_kwargs))
            if act is not None:
                layers.append(ACTIVATIONS[act])
        super().__init__(*layers)

        # everything needed to rebuild the net (transfer learning), as in the reference
        conv_kwargs.update(dict(in_channels=in_channels, out_channels=out_channels,
                                kernel_size=kernel_size, conv_dim=conv_dim, hidden_sizes=hidden_sizes,
                                acts=acts, pre_act=pre_act))
        self.conv_kwargs = conv_kwargs
        self._acts = tuple(acts)
        self._pre_act = pre_act

    # ---- fused execution ------------------------------------------------------------
    @property
    def fusable(self):
        """True when the whole stack runs as fused conv+activation kernels under one
        autograd node (activations the kernel knows, none on the input or output side
        that it cannot differentiate)."""
        return (self._pre_act is None and all(a in _FUSED_ACTS for a in self._acts)
                and self._acts[-1] in (None, 'none'))

    def _convs(self):
        return [m for m in self if hasattr(m, 'standard_weight')]

    @property
    def fused2d_ok(self):
        """True for the shape the single-kernel 2-D coupling step is specialised for:
        ConvAct(1 -> 8 -> 8 -> P), 3x3, tanh, tanh, none."""
        kw = self.conv_kwargs
        return (self._pre_act is None and kw['conv_dim'] == 2 and kw['kernel_size'] == 3
                and kw['in_channels'] == 1 and list(kw['hidden_sizes']) == [8, 8]
                and tuple(self._acts) == ('tanh', 'tanh', None)
                and kw.get('padding', 'same') == 'same' and kw.get('padding_mode') == 'circular'
                and all(k not in kw for k in ('stride', 'dilation', 'groups')))

    @property
    def fusednd_ok(self):
        """The conditioner shape of the N-D tensor-core step: ConvAct(1 -> H -> H -> P) with H in {8, 16, 32, 64},
        3^D taps, tanh, tanh, none, on a 2-D, 3-D or 4-D lattice (Conv2d / Conv3d / Conv4d layers).  (The 2-D, H = 8
        case is also `fused2d_ok` and takes the single-kernel step.)"""
        kw = self.conv_kwargs
        hidden = list(kw['hidden_sizes'])
        return (self._pre_act is None and kw['conv_dim'] in (2, 3, 4) and kw['kernel_size'] == 3
                and kw['in_channels'] == 1 and len(hidden) == 2 and hidden[0] == hidden[1]
                and hidden[0] in (8, 16, 32, 64)
                and tuple(self._acts) == ('tanh', 'tanh', None)
                and kw.get('padding', 'same') == 'same' and kw.get('padding_mode') == 'circular'
                and all(k not in kw for k in ('stride', 'dilation', 'groups')))

    def forward_masked(self, x, mask, keep):
        """Conditioner output for a field x (B, *L) of which only the sites with
        mask == keep are visible (the frozen partition): Mask.split fused into the
        first conv layer.  Needs in_channels == 1."""
        convs = self._convs()
        if convs[0].in_channels != 1:
            raise ValueError("forward_masked expects a single input channel")
        return _ops.conv_stack(x.unsqueeze(1), [c.standard_weight() for c in convs], [c.bias for c in convs],
                               self._acts, convs[0].kernel_size[0], in_mask=mask, in_keep=keep)

    def forward(self, inp):
        if self.fusable:
            convs = self._convs()
            return _ops.conv_stack(inp, [c.standard_weight() for c in convs], [c.bias for c in convs],
                                   self._acts, convs[0].kernel_size[0])
        return super().forward(inp)

    def set_param2zero(self):
        for p in self.parameters():
            torch.nn.init.zeros_(p)

    def transfer(self, scale_factor=1, **extra):
        if scale_factor != 1:
            raise NotImplementedError("kernel rescaling on transfer is not implemented "
                                      "(marked outdated in the reference as well)")
        return copy.deepcopy(self)


class SplineNet(torch.nn.Module):
    """Trainable knots of ONE shared 1-D rational-quadratic spline
    (reference modules.py:276-391, spline_shape=[] only).

    The first knot sits at (xlim[0], ylim[0]) and the last at (xlim[1], ylim[1]); the
    K-1 widths / heights in between are softmax(weights_x / weights_y) and the K knot
    derivatives are softplus_{beta=ln2}(weights_d) -- or, with smooth=True, the average
    slope of the adjacent segments.
    """

    def __init__(self, knots_len, xlim=(0, 1), ylim=(0, 1), knots_x=None, knots_y=None, knots_d=None,
                 spline_shape=[], knots_axis=-1, smooth=False, Spline=RQSpline, label='spline',
                 **spline_kwargs):
        super().__init__()
        if len(spline_shape) > 0:
            raise NotImplementedError("SplineNet: only one shared spline (spline_shape=[]) is supported")
        if knots_x is not None or knots_y is not None or knots_d is not None:
            raise NotImplementedError("SplineNet with fixed knots is outside the accelerated hot path")
        if knots_len < 2:
            raise AssertionError("oops: knots_len < 2 for splines")
        self.label = label
        self.knots_len = knots_len
        self.knots_x = self.knots_y = self.knots_d = None
        self.spline_shape, self.knots_axis = spline_shape, knots_axis
        self.Spline, self.spline_kwargs = Spline, spline_kwargs
        self.smooth = smooth
        self.xlim, self.xwidth = xlim, xlim[1] - xlim[0]
        self.ylim, self.ywidth = ylim, ylim[1] - ylim[0]
        self.weights_x = torch.nn.Parameter(torch.zeros(knots_len - 1))
        self.weights_y = torch.nn.Parameter(torch.zeros(knots_len - 1))
        self.weights_d = None if smooth else torch.nn.Parameter(torch.zeros(knots_len))

    def knots(self, both_ends=False):
        """Differentiable float32 knot table [3, K] = knots_x | knots_y | knots_d, or with
        both_ends=True [5, K] with xlim[1] - knots_x and ylim[1] - knots_y appended.

        Every row is built from cumulative softmax sums on its own side: knots_x from
        the left, its complement from the right (so a knot close to xlim[1] is known to
        the relative precision of its distance from xlim[1], not of its value), and the
        bin widths / heights used for the smooth derivatives are the softmax terms
        themselves.  The end knots are exactly xlim / ylim.  One kernel (`nfk_knots_*`)."""
        table = _ops.knot_table(self.weights_x, self.weights_y, self.weights_d, self.xlim, self.ylim)
        return table if both_ends else table[:3]

    def make_spline(self):
        table = self.knots()
        spline = self.Spline(knots_x=table[0], knots_y=table[1], knots_d=table[2], **self.spline_kwargs)
        spline._packed = table        # the kernel's input as it stands: no re-stacking of the rows
        return spline

    def forward(self, x):
        return self.make_spline()(x.ravel()).reshape(x.shape)

    def backward(self, x):
        return self.make_spline().backward(x.ravel()).reshape(x.shape)
