"""Multi-GPU plumbing: one process per GPU, batch data parallelism over NCCL
(reference src/device/_core.py).

The reference wraps `net_` in DistributedDataParallel, whose reducer walks gradient
buckets on a side stream.  All parameters of a flow fit in well under a megabyte, so
here the gradients of every parameter are views into ONE flat float32 buffer and a
training step issues a single in-place all-reduce (average) on the compute stream,
right after `backward()`.  Sampling needs no collective; diagnostics use one
all-gather.  Samples are never split across devices.
"""

import os
import warnings
from functools import partial

import torch
import torch.distributed as dist


class ModelDeviceHandler:
    """Keeps track of rank / world size for a Model and moves it between devices
    (reference ModelDeviceHandler, device/_core.py:27-95)."""

    def __init__(self, model):
        self._model = model
        self.nranks = 1
        self.rank = 0
        self._flat_grad = None
        self._flat_params = None

    def to(self, *args, **kwargs):
        self._model.net_.to(*args, **kwargs)
        self._model.prior.to(*args, **kwargs)

    # ---- data-parallel set-up ------------------------------------------------------
    def ddp_wrapper(self, rank, nranks, device=None):
        """Place the model on this rank's device and prepare gradient averaging.
        Keeps the reference's name; there is no wrapper object: `net_` stays the plain
        ModuleList_, so attribute access needs no pass-through shim (device/_core.py:15-23)."""
        if device is None:
            device = torch.device('cuda', rank) if torch.cuda.is_available() else torch.device('cpu')
        if device.type == 'cuda':
            torch.cuda.set_device(device)
        self._model.prior.to(device=device)
        self._model.net_.to(device=device)
        self.nranks, self.rank = nranks, rank
        if nranks > 1:
            self.broadcast_parameters()
            self.flatten_gradients()

    def broadcast_parameters(self, src=0):
        """Every rank starts from rank `src`'s weights (what DDP does at wrap time)."""
        for p in self._model.net_.parameters():
            dist.broadcast(p.data, src=src)

    def flatten_gradients(self):
        """Make every parameter's .grad a view into one contiguous buffer."""
        params = [p for p in self._model.net_.parameters() if p.requires_grad]
        if not params:
            return
        total = sum(p.numel() for p in params)
        flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
        offset = 0
        for p in params:
            n = p.numel()
            p.grad = flat[offset:offset + n].view_as(p)
            offset += n
        self._flat_grad, self._flat_params = flat, params

    def zero_grad(self, optimizer=None):
        """Clear gradients.  With a flat buffer they are zeroed in place (setting them to
        None, torch's default, would detach the views)."""
        if self._flat_grad is not None:
            self._flat_grad.zero_()
        elif optimizer is not None:
            optimizer.zero_grad()

    def sync_gradients(self):
        """Average gradients over ranks: one all-reduce of the flat buffer, enqueued on
        the current stream (reference: DDP bucketed all-reduce, device/_core.py:46)."""
        if self.nranks == 1:
            return
        if self._flat_grad is None:
            self.flatten_gradients()
        self._reattach()
        if dist.get_backend() == 'nccl':
            dist.all_reduce(self._flat_grad, op=dist.ReduceOp.AVG)     # ncclAvg: one launch
        else:                                                          # gloo (CPU tests) has no AVG
            dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM)
            self._flat_grad.div_(self.nranks)

    def _reattach(self):
        """Autograd accumulates in place into an existing .grad, so the views normally
        survive; if something replaced one (or set it to None) fold it back in."""
        offset = 0
        for p in self._flat_params:
            n = p.numel()
            view = self._flat_grad[offset:offset + n].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            offset += n

    def all_reduce_mean(self, x):
        """Mean of a tensor over ranks (loss / acceptance statistics)."""
        if self.nranks == 1:
            return x
        y = x.clone()
        if dist.get_backend() == 'nccl':
            dist.all_reduce(y, op=dist.ReduceOp.AVG)
            return y
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y / self.nranks

    def all_gather_into_tensor(self, x):
        """Concatenate x from every rank along axis 0 (device/_core.py:87-95)."""
        if self.nranks == 1:
            return x
        shape = list(x.shape)
        shape[0] *= self.nranks
        out = torch.zeros(*shape, dtype=x.dtype, device=x.device)
        if dist.get_backend() == 'gloo':
            parts = [torch.empty_like(x) for _ in range(self.nranks)]
            dist.all_gather(parts, x.contiguous())
            return torch.cat(parts, dim=0)
        dist.all_gather_into_tensor(out, x.contiguous())
        return out

    # ---- process spawning ----------------------------------------------------------
    def spawnprocesses(self, fn, nranks, master_port=12354, seeds_torch=None, *args, **kwargs):
        """Run fn(model, *args, **kwargs) in `nranks` processes, one per GPU
        (reference spawnprocesses, device/_core.py:51-85)."""
        seeds_torch = prepare_seeds(nranks, seeds_torch)
        wrapped = DistributedFunc(fn)
        try:
            torch.multiprocessing.spawn(partial(wrapped, **kwargs),
                                        args=(nranks, master_port, seeds_torch, self._model) + tuple(args),
                                        nprocs=nranks, join=True)
        except torch.multiprocessing.ProcessException as err:
            warnings.warn("Distributed run could not be spawned. If the default master port is taken, "
                          "choose another one (master_port=...).")
            raise err


class DistributedFunc:
    """Child-process entry: join the process group, place the model, seed, run, leave."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, rank, nranks, master_port, seeds_torch, model, *args, **kwargs):
        setup_process_group(rank, nranks, master_port=master_port)
        try:
            model.device_handler.ddp_wrapper(rank, nranks)
            torch.manual_seed(seeds_torch[rank])     # independent Philox key per rank
            return self.fn(model, *args, **kwargs)
        finally:
            dist.destroy_process_group()


def setup_process_group(rank, world_size, master_addr='127.0.0.1', master_port=12354, backend=None):
    """NCCL (GPU) or gloo (CPU) process group on this node (device/_core.py:120-133)."""
    os.environ['MASTER_ADDR'] = master_addr
    os.environ['MASTER_PORT'] = str(master_port)
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(rank)
    dist.init_process_group(backend=backend, rank=rank, world_size=world_size)


def prepare_seeds(nranks, seeds_torch):
    """One torch seed per rank; random ones when none are given (device/_core.py:136-158)."""
    if seeds_torch is None:
        return gen_seed(size=(nranks,))
    if len(seeds_torch) != nranks:
        raise AssertionError("Numbers of seeds != nranks")
    return list(seeds_torch)


def gen_seed(size=None):
    """Random seed(s) below 2**32 - 1 (numpy's limit): a number, or a list for a size."""
    draw = torch.randint(2 ** 32 - 1, size=[1] if size is None else size, device='cpu').tolist()
    return draw[0] if size is None else draw
