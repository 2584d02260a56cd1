"""Prior distributions (reference src/prior/prior.py).

`NormalPrior.sample_` draws the batch AND its log-density in one kernel: Philox4x32-10
+ Box-Muller writes x once and a block reduction yields logr[B] (the reference makes
one sampling pass and ~5 elementwise passes, prior.py:23-36).
"""

import numpy as np
import torch

from .. import _ops


class Prior:
    """Interface of a prior: sample, sample_ (with log-density), log_prob, to, shape, nvar."""

    propagate_density = False

    def sample(self, batch_size=1):
        raise NotImplementedError

    def sample_(self, batch_size=1):
        raise NotImplementedError

    def log_prob(self, x):
        raise NotImplementedError

    @staticmethod
    def manual_seed(seed):
        """Same convention as the reference (prior.py:38-41): an int seeds torch's
        default generator, which is where the Philox key is taken from."""
        if isinstance(seed, int):
            torch.manual_seed(seed)

    @property
    def nvar(self):
        return int(np.prod(self.shape))


class NormalPrior(Prior):
    """Independent normal variables with site-wise `loc` and `scale`
    (reference NormalPrior, prior.py:92-125).  Give either `shape` (standard normal)
    or `loc` and `scale` tensors of the lattice shape."""

    def __init__(self, loc=None, scale=None, shape=None, seed=None, **kwargs):
        if shape is not None:
            shape = (shape,) if isinstance(shape, int) else tuple(shape)
            self._standard = True
            loc, scale = torch.zeros(shape), torch.ones(shape)
        else:
            if loc is None or scale is None:
                raise ValueError("NormalPrior needs `shape`, or `loc` and `scale`")
            self._standard = False
            loc, scale = torch.as_tensor(loc), torch.as_tensor(scale)
            shape = tuple(loc.shape)
        self.shape = shape
        self._loc = loc.to(torch.float32).contiguous()
        self._scale = scale.to(torch.float32).contiguous()
        self._calls = 0          # Philox stream offset: one stream per draw
        self._block_calls = 0    # the same for the block proposals of setup_blockupdater (persistent across calls)
        self._state = None       # device-resident {seed, offset}: see use_device_state()
        Prior.manual_seed(seed)

    def use_device_state(self, enable=True):
        """Keep the generator state (Philox key, stream offset) in device memory, advanced by every
        draw on the stream.  Needed when draws are captured in a CUDA graph (Fitter's graph mode):
        host-side offsets would be frozen into the graph and every replay would repeat the batch."""
        if not enable:
            self._state = None
        elif self._state is None:
            if not self._loc.is_cuda:
                raise RuntimeError("NormalPrior.use_device_state: move the prior to CUDA first")
            seed = torch.initial_seed() & (2 ** 63 - 1)
            self._state = torch.tensor([seed, self._calls + 1], dtype=torch.int64, device=self._loc.device)

    # the reference exposes the torch distribution object as `.dist`
    @property
    def dist(self):
        return torch.distributions.normal.Normal(self._loc, self._scale)

    def _args(self):
        if self._standard:
            return None, None
        return self._loc, self._scale

    def _draw(self, batch_size, with_logprob):
        if not self._loc.is_cuda:
            raise RuntimeError("NormalPrior: sampling runs on CUDA only; move the prior with "
                               ".to('cuda') (normflow__b200 has no CPU fallback)")
        loc, scale = self._args()
        self._calls += 1
        if self._state is not None:
            return _ops.prior_sample(int(batch_size), self.shape, loc, scale, 0, 0, self._loc.device,
                                     with_logprob=with_logprob, state=self._state)
        return _ops.prior_sample(int(batch_size), self.shape, loc, scale,
                                 seed=torch.initial_seed(), offset=self._calls,
                                 device=self._loc.device, with_logprob=with_logprob)

    def sample(self, batch_size=1):
        return self._draw(batch_size, False)[0]

    def sample_(self, batch_size=1):
        """(x, log r(x)) -- reference prior.py:26-28."""
        return self._draw(batch_size, True)

    def log_prob(self, x):
        """sum over sites of the normal log-density (reference prior.py:30-36)."""
        if self.propagate_density:
            raise NotImplementedError("propagate_density is not part of the accelerated path")
        loc, scale = self._args()
        return _ops.prior_logprob(x, loc, scale)

    def setup_blockupdater(self, block_len):
        """Attach `self.blockupdater`, which redraws one block of `block_len` consecutive variables
        of a batch in place (prior.py:106-112; like the reference it takes loc / scale of the first
        block for every block).  Its draws come from a Philox stream range of their own, far from
        the one `sample` walks, so proposals never repeat values this prior has already produced."""
        old = getattr(self, 'blockupdater', None)
        if old is not None:
            # the block stream keeps its place: a new updater (or a repeated call with the same block
            # length, which is what BlockedMCMCSampler.sample__ does every call) continues after the last
            # block draw instead of replaying the sequence of proposals from its start
            self._block_calls = old.chopped_prior._calls - (1 << 40)
            if old.block_len == block_len and old.chopped_prior._loc.device == self._loc.device:
                return
        chopped = NormalPrior(loc=self._loc.reshape(-1)[:block_len], scale=self._scale.reshape(-1)[:block_len])
        chopped._standard = self._standard
        chopped._calls = (1 << 40) + max(self._block_calls, self._calls)
        self.blockupdater = BlockUpdater(chopped, block_len)

    def to(self, *args, **kwargs):
        """Move loc/scale; samples are created on the same device (prior.py:110-116).
        The kernels are float32, so a dtype request other than float32 is refused."""
        dtype = kwargs.get('dtype', None)
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
        if dtype not in (None, torch.float32):
            raise TypeError("normflow__b200 runs in float32 only")
        kwargs = {k: v for k, v in kwargs.items() if not (k == 'dtype' and v is None)}
        self._loc = self._loc.to(*args, **kwargs)
        self._scale = self._scale.to(*args, **kwargs)
        if self._state is not None and self._state.device != self._loc.device:
            self._state = self._state.to(self._loc.device) if self._loc.is_cuda else None

    @property
    def parameters(self):
        return dict(loc=self._loc, scale=self._scale)


class BlockUpdater:
    """In-place redraw / restore of block `block_ind` of every sample of a batch, the field seen
    as (B, n_blocks, block_len) (reference BlockUpdater, prior.py:161-178)."""

    def __init__(self, chopped_prior, block_len):
        self.block_len = block_len
        self.chopped_prior = chopped_prior
        self.backup_block = None

    def _blocks(self, x):
        return x.view(x.shape[0], -1, self.block_len)

    def __call__(self, x, block_ind):
        blocks = self._blocks(x)
        self.backup_block = blocks[:, block_ind].clone()
        blocks[:, block_ind] = self.chopped_prior.sample(x.shape[0])

    def restore(self, x, block_ind, restore_ind=slice(None)):
        self._blocks(x)[restore_ind, block_ind] = self.backup_block[restore_ind]
