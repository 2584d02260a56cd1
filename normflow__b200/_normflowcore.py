"""Model, Posterior and Fitter: the drivers of the hot loop
(reference src/_normflowcore.py).

    model = Model(prior=..., net_=..., action=...)
    model.fit(n_epochs=..., batch_size=...)          # training      (Fitter)
    y = model.posterior.sample(batch_size)           # raw proposals (Posterior)
    y = model.mcmc.sample(batch_size)                # with Metropolis accept/reject

The classes only sequence the work: every pass over a batch of lattice fields (prior
draw + log-density, coupling layers with log|det J|, action, accept/reject, and their
gradients) is a kernel of libnormflow_b200.so reached through prior / net_ / action.
"""

import os
import time

import numpy as np
import torch

from .mcmc import MCMCSampler, BlockedMCMCSampler
from .lib.combo import estimate_logz, fmt_val_err
from .device import ModelDeviceHandler


class Model:
    """Bundle of a prior, an invertible network and an action
    (reference Model, _normflowcore.py:33-67).

    prior  : e.g. prior.NormalPrior(shape=lattice_shape)
    net_   : nn.ModuleList_ (or any layer with forward/backward returning log-Jacobians)
    action : e.g. action.ScalarPhi4Action(...)
    """

    def __init__(self, *, prior, net_, action, name=None):
        self.name = name
        self.net_ = net_
        self.prior = prior
        self.action = action

        self.fit = Fitter(self)
        self.posterior = Posterior(self)
        self.raw_dist = self.posterior      # older alias kept by the reference
        self.mcmc = MCMCSampler(self)
        self.blocked_mcmc = BlockedMCMCSampler(self)
        self.device_handler = ModelDeviceHandler(self)

    def transform(self, x):
        return self.net_(x)[0]


class Posterior:
    """Samples drawn straight from the flow, no accept/reject
    (reference Posterior, _normflowcore.py:70-119)."""

    def __init__(self, model):
        self._model = model
        # Opt-in (launch-bound lattices): replay prior draw + flow + action of one batch size as ONE captured
        # CUDA graph.  The returned tensors are the graph's output buffers, overwritten by the next call with the
        # same batch size -- clone what must outlive it.
        self.cuda_graph = False
        self._graphs = {}

    def _graph_key(self, batch_size):
        params = tuple(p.data_ptr() for p in self._model.net_.parameters())
        return (int(batch_size), params, id(self._model.prior), id(self._model.action))

    def _sample_graph(self, batch_size):
        """(y, logq, logp) of a fresh batch through a CUDA graph captured on first use of this batch size (and
        re-captured if the parameters have been re-allocated).  Parameter VALUES are read at replay, so training
        between calls is seen; the prior's generator state lives on the device and advances inside the graph."""
        key = self._graph_key(batch_size)
        entry = self._graphs.get(batch_size)
        if entry is None or entry[0] != key:
            prior = self._model.prior
            prior.use_device_state()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                   # first-use initialisation outside the capture
                    self._sample_eager(batch_size)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._sample_eager(batch_size)
            entry = (key, graph, out)
            self._graphs[batch_size] = entry
        entry[1].replay()
        return entry[2]

    def _sample_eager(self, batch_size):
        x, logr = self._model.prior.sample_(batch_size)
        y, logJ = self._model.net_(x)
        return y, logr - logJ, -self._model.action(y)

    def _graph_ok(self, kwargs):
        if not self.cuda_graph or kwargs.get('preprocess_func') is not None:
            return False
        params = list(self._model.net_.parameters())
        return bool(params) and all(p.is_cuda for p in params) and hasattr(self._model.prior, 'use_device_state')

    @torch.no_grad()
    def sample(self, batch_size=1, **kwargs):
        return self.sample_(batch_size=batch_size, **kwargs)[0]

    @torch.no_grad()
    def sample_(self, batch_size=1, preprocess_func=None):
        """(y, log q(y)); `preprocess_func(x, logr)` may edit the prior draw first."""
        if self._graph_ok(dict(preprocess_func=preprocess_func)):
            return self._sample_graph(batch_size)[:2]
        x, logr = self._model.prior.sample_(batch_size)
        if preprocess_func is not None:
            x, logr = preprocess_func(x, logr)
        y, logJ = self._model.net_(x)
        return y, logr - logJ

    @torch.no_grad()
    def sample__(self, batch_size=1, **kwargs):
        """(y, log q(y), log p(y) + log z): the benchmark's "flow fwd + logJ + action"."""
        if self._graph_ok(kwargs):
            return self._sample_graph(batch_size)
        y, logq = self.sample_(batch_size=batch_size, **kwargs)
        return y, logq, -self._model.action(y)

    @torch.no_grad()
    def log_prob(self, y):
        """log q(y) through the inverse flow."""
        x, minus_logJ = self._model.net_.backward(y)
        return self._model.prior.log_prob(x) + minus_logJ


class Fitter:
    """Trains the model by minimising a divergence between q and p
    (reference Fitter, _normflowcore.py:122-428)."""

    def __init__(self, model):
        self._model = model
        self.train_batch_size = 1
        self.train_history = dict(loss=[], logqp=[], logz=[], ess=[], rho=[], accept_rate=[])
        self.hyperparam = dict(lr=0.001, weight_decay=0.01)
        self.checkpoint_dict = dict(display=False, print_stride=100, print_batch_size=1024,
                                    print_extra_func=None, snapshot_path=None, epochs_run=0)
        # Opt-in: replay the whole optimisation step (draw, flow, action, backward, optimiser) as ONE
        # captured CUDA graph.  For small lattices the step is pure launch latency (tens of kernels of
        # a few microseconds each); see _train_graph.  Also switched on by NFK_CUDA_GRAPH=1.
        self.cuda_graph = os.environ.get('NFK_CUDA_GRAPH') == '1'
        self.graph_timing = None          # {'replays', 'ms'} of the last graph-mode fit (device time of the replays)
        self._pending_losses = []     # device scalars of the epochs since the last flush
        self._guard = None

    def __call__(self, n_epochs=1000, save_every=None, batch_size=64,
                 optimizer_class=torch.optim.AdamW, scheduler=None, loss_fn=None,
                 hyperparam={}, checkpoint_dict={}):
        """Train for `n_epochs` steps of `batch_size` fresh samples each.

        save_every      : snapshot period in epochs (default: only at the end)
        optimizer_class : torch optimiser class (AdamW by default)
        scheduler       : callable optimiser -> lr scheduler, or None
        loss_fn         : f(logq, logp) -> scalar; default reverse KL (`calc_kl_mean`)
        hyperparam      : optimiser keyword arguments (lr, weight_decay, ...)
        checkpoint_dict : print_stride, print_batch_size, snapshot_path, ...
        """
        self.hyperparam.update(hyperparam)
        self.checkpoint_dict.update(checkpoint_dict)
        snapshot_path = self.checkpoint_dict['snapshot_path']
        if save_every is None:
            save_every = n_epochs

        if snapshot_path is None:
            print("Not saving model snapshots")
        elif os.path.exists(snapshot_path):
            print(f"Trying to load snapshot from {snapshot_path}")
            self._load_snapshot()
        else:
            print("Starting training from scratch")

        self.loss_fn = Fitter.calc_kl_mean if loss_fn is None else loss_fn

        # (the reference's test for parameter groups, `'_groups' is net_.__dict__.keys()`,
        # is never true, so it always optimises net_.parameters(); kept that way)
        hyper = dict(self.hyperparam)
        self._graph_ok = self._graph_mode_possible(optimizer_class, scheduler, hyper)
        params = list(self._model.net_.parameters())
        on_cuda = bool(params) and all(p.is_cuda for p in params)
        if self._graph_ok:
            hyper['fused'] = True       # one kernel over all parameters; honours the device-side NaN guard
            hyper['capturable'] = True  # step counters live on the device
        elif (on_cuda and optimizer_class in (torch.optim.AdamW, torch.optim.Adam)
              and 'fused' not in hyper and not hyper.get('foreach')):
            hyper['fused'] = True       # eager default on CUDA: one launch instead of a foreach chain
        self.optimizer = optimizer_class(params, **hyper)
        self.scheduler = None if scheduler is None else scheduler(self.optimizer)
        # Device-side divergence guard (the reference's host-side `if isnan(loss)`, :289): the fused optimisers
        # skip their update when `found_inf` is set, so neither the eager nor the captured step waits for the
        # GPU.  Any other optimiser keeps the host-side test.
        self._guard = None
        if hyper.get('fused') and on_cuda:
            dev = params[0].device
            self._guard = dict(found_inf=torch.zeros((), dtype=torch.float32, device=dev),
                               n_skipped=torch.zeros((), dtype=torch.float32, device=dev))
            self.optimizer.found_inf = self._guard['found_inf']
        return self.train(n_epochs, batch_size, save_every)

    def _graph_mode_possible(self, optimizer_class, scheduler, hyper):
        """Graph mode needs: the switch, CUDA parameters, no scheduler (a captured step has its learning
        rate baked in) and an optimiser with a fused, capturable implementation (`fused=True` in the
        hyperparameters is what graph mode sets anyway; only an explicit False declines).  With several
        ranks the gradient all-reduce is captured with the step (NCCL is capturable)."""
        if not self.cuda_graph:
            return False
        why = None
        params = list(self._model.net_.parameters())
        if scheduler is not None:
            why = "a learning-rate scheduler is set (a captured step has its learning rate baked in)"
        elif optimizer_class not in (torch.optim.AdamW, torch.optim.Adam):
            why = f"{optimizer_class.__name__} has no fused, capturable implementation"
        elif hyper.get('fused', True) is False or hyper.get('capturable', True) is False:
            why = "hyperparam asks for fused=False / capturable=False"
        elif not params or not all(p.is_cuda for p in params):
            why = "the parameters are not on a CUDA device"
        if why is not None:
            if self._model.device_handler.rank == 0:
                print(f"fit: cuda_graph requested but declined -- {why}; training eagerly")
            return False
        return True

    # ---- snapshots ({"MODEL_STATE", "EPOCHS_RUN"}, reference :221-247) -------------
    def _load_snapshot(self):
        path = self.checkpoint_dict['snapshot_path']
        if torch.cuda.is_available():
            loc = f"cuda:{self._model.device_handler.rank}"
            print(f"GPU: Attempting to load saved model into {loc}")
        else:
            loc = None
            print("CPU: Attempting to load saved model")
        snapshot = torch.load(path, map_location=loc)
        state = {k: (v.float() if torch.is_floating_point(v) else v)
                 for k, v in snapshot["MODEL_STATE"].items()}     # reference snapshots are float64
        self._model.net_.load_state_dict(state)
        self.checkpoint_dict['epochs_run'] = snapshot['EPOCHS_RUN']
        print(f"Snapshot found: {path}\nResuming training via Saved Snapshot at Epoch {snapshot['EPOCHS_RUN']}")

    def _save_snapshot(self, epoch):
        path = self.checkpoint_dict['snapshot_path']
        epochs_run = epoch + self.checkpoint_dict['epochs_run']
        new_path = path.rsplit('.', 2)[0] + ".E" + str(epochs_run) + ".tar"
        torch.save({"MODEL_STATE": self._model.net_.state_dict(), "EPOCHS_RUN": epochs_run}, new_path)
        print(f"Epoch {epochs_run} | Model Snapshot saved at {new_path}")

    # ---- the loop -------------------------------------------------------------------
    def train(self, n_epochs, batch_size, save_every):
        """The epoch loop; call through `__call__` (which builds the optimiser) at least once."""
        self.train_batch_size = batch_size
        t_start = time.time()
        loss = None
        n_eager = n_epochs
        if getattr(self, '_graph_ok', False) and n_epochs > self._GRAPH_WARMUP:
            n_eager = self._GRAPH_WARMUP          # the first steps run eagerly: they are the graph's warm-up
        for epoch in range(1, n_eager + 1):
            loss = self.step()[0]
            self.checkpoint(epoch, loss, save_every)
            if self.scheduler is not None:
                self.scheduler.step()
        if n_eager < n_epochs:
            # drop the eager autograd graph first: its AccumulateGrad nodes are tied to the stream the
            # eager steps ran on, and a capture may not depend on the legacy default stream
            loss = None
            self._flush_losses()
            loss = self._train_graph(n_eager + 1, n_epochs, save_every)
        self._flush_losses()
        self._report_skipped()
        if n_epochs > 0 and self._model.device_handler.rank == 0:
            print(f"({loss.device}) Time = {time.time() - t_start:.3g} sec.")

    _GRAPH_WARMUP = 3

    def _set_found_inf(self, loss):
        """found_inf <- 1 when this step must not be taken, evaluated on the device.  One rank: the loss is not
        finite (the reference's test, :289).  Several ranks: the AVERAGED gradient is not finite -- a loss that
        diverged on any rank poisons it on every rank, so all ranks skip the same steps and stay in step."""
        guard, handler = self._guard, self._model.device_handler
        probe = loss.detach() if handler.nranks == 1 or handler._flat_grad is None else handler._flat_grad.sum()
        guard['found_inf'].copy_(1.0 - torch.isfinite(probe).to(torch.float32))
        guard['n_skipped'].add_(guard['found_inf'])

    def _report_skipped(self):
        if getattr(self, '_guard', None) is None:
            return
        skipped = int(self._guard['n_skipped'].item())
        if skipped:
            print(f"OOPS: loss was divergent in {skipped} step(s) -> no *step* was taken there.")
            self._guard['n_skipped'].zero_()

    def _train_graph(self, first_epoch, n_epochs, save_every):
        """Epochs first_epoch..n_epochs as replays of one captured CUDA graph of `step`.

        What makes the step capturable: every kernel of this package is enqueued on the current
        stream without host synchronisation; the prior keeps its Philox state on the device
        (`NormalPrior.use_device_state`), so each replay draws a fresh batch; the reference's
        host-side `if isnan(loss)` becomes the fused optimiser's `found_inf` flag (the update is
        skipped on the device); the loss of every epoch goes to a device buffer that is read back
        only when something is printed or saved.  With several ranks the flat-gradient all-reduce
        (NCCL) is part of the captured step."""
        model, handler = self._model, self._model.device_handler
        dev = next(model.net_.parameters()).device
        model.prior.use_device_state(True)
        n_graph = n_epochs - first_epoch + 1
        loss_buf = torch.zeros(n_graph, dtype=torch.float32, device=dev)
        slot = torch.zeros(1, dtype=torch.int64, device=dev)
        multi = handler.nranks > 1

        def body():
            x, logr = model.prior.sample_(self.train_batch_size)
            y, logJ = model.net_(x)
            loss = self.loss_fn(logr - logJ, -model.action(y))
            if multi:
                handler.zero_grad()                     # the flat buffer, in place
            loss.backward()
            handler.sync_gradients()
            self._set_found_inf(loss)
            self.optimizer.step()                       # no-op on the device when found_inf is set
            loss_buf.index_copy_(0, slot, loss.detach().reshape(1))
            slot.add_(1)
            return loss

        if not multi:
            self.optimizer.zero_grad(set_to_none=True)  # gradients get allocated inside the graph's pool
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        # thread-local capture mode: NCCL's watchdog thread may query events while this thread captures
        with torch.cuda.graph(graph, capture_error_mode="thread_local" if multi else "global"):
            static_loss = body()
        # (capturing does not execute: epoch `first_epoch` is the first replay)
        flushed = 0

        def flush(upto):                                # device losses -> train_history (one sync)
            nonlocal flushed
            if upto > flushed and handler.rank == 0:
                self.train_history['loss'].extend(loss_buf[flushed:upto].tolist())
            flushed = max(flushed, upto)

        print_stride = self.checkpoint_dict['print_stride']
        snapshot_path = self.checkpoint_dict['snapshot_path']
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for epoch in range(first_epoch, n_epochs + 1):
            graph.replay()
            diag = epoch == 1 or epoch == 10 or epoch % print_stride == 0
            save = snapshot_path is not None and epoch % save_every == 0
            if diag or save:
                flush(epoch - first_epoch + 1)
                self._checkpoint_tail(epoch, save_every)
        t1.record()
        flush(n_graph)
        t1.synchronize()
        # device time of the replayed epochs, capture excluded (diagnostics: bench.py reads it)
        self.graph_timing = {'replays': n_graph, 'ms': t0.elapsed_time(t1)}
        return static_loss.detach()

    def step(self):
        """One optimisation step on a fresh batch (reference Fitter.step, :275-294).  No host
        synchronisation when the optimiser is a fused Adam / AdamW (the default on CUDA)."""
        model, handler = self._model, self._model.device_handler
        x, logr = model.prior.sample_(self.train_batch_size)
        y, logJ = model.net_(x)
        logq = logr - logJ
        logp = -model.action(y)
        loss = self.loss_fn(logq, logp)

        handler.zero_grad(self.optimizer)
        loss.backward()
        handler.sync_gradients()          # one flat all-reduce when nranks > 1
        if getattr(self, '_guard', None) is not None:
            self._set_found_inf(loss)
            self.optimizer.step()         # skipped on the device when the guard is set
        elif torch.isnan(loss):
            print("OOPS: loss is divergent -> no *step* is taken.")
        else:
            self.optimizer.step()
        return loss, logq - logp

    def checkpoint(self, epoch, loss, save_every):
        """Book-keeping of an eager epoch.  The loss stays a device scalar until something is printed or
        saved (or training ends): `train_history['loss']` is filled in batches, without a sync per epoch
        (the reference calls `loss.item()` every epoch, :305)."""
        if self._model.device_handler.rank == 0:
            self._pending_losses.append(loss.detach())
        print_stride = self.checkpoint_dict['print_stride']
        snapshot_path = self.checkpoint_dict['snapshot_path']
        diag = epoch == 1 or epoch == 10 or epoch % print_stride == 0
        if diag or (snapshot_path is not None and epoch % save_every == 0) or len(self._pending_losses) >= 4096:
            self._flush_losses()
        self._checkpoint_tail(epoch, save_every)

    def _flush_losses(self):
        pending = self._pending_losses
        if pending:
            self.train_history['loss'].extend(torch.stack(pending).tolist())
            pending.clear()

    def _checkpoint_tail(self, epoch, save_every):
        """Snapshot and diagnostics of an epoch whose loss is already in train_history."""
        handler = self._model.device_handler
        rank = handler.rank
        print_stride = self.checkpoint_dict['print_stride']
        print_batch_size = self.checkpoint_dict['print_batch_size'] // handler.nranks
        snapshot_path = self.checkpoint_dict['snapshot_path']
        if rank == 0 and snapshot_path is not None and (epoch % save_every == 0):
            self._save_snapshot(epoch)

        if epoch == 1 or epoch == 10 or (epoch % print_stride == 0):
            _, logq, logp = self._model.posterior.sample__(print_batch_size)
            logq = handler.all_gather_into_tensor(logq)
            logp = handler.all_gather_into_tensor(logp)
            if rank == 0:
                loss_ = self.loss_fn(logq, logp)
                self._append_to_train_history(logq, logp)
                self.print_fit_status(epoch, loss=loss_)

    # ---- losses and diagnostics on [B] vectors (negligible bytes: tensor expressions) --
    @staticmethod
    def calc_kl_mean(logq, logp):
        """Reverse KL up to log z, estimated with samples from q."""
        return (logq - logp).mean()

    @staticmethod
    def calc_kl_var(logq, logp):
        return (logq - logp).var()

    @staticmethod
    def calc_corrcoef(logq, logp):
        return torch.corrcoef(torch.stack([logq, logp]))[0, 1]

    @staticmethod
    def calc_direct_kl_mean(logq, logp):
        r"""sum(p/q (log(p/q) + log z)) / sum(p/q) with log z = log mean(p/q);
        invariant under rescaling p or q."""
        logpq = logp - logq
        logpq = logpq - (torch.logsumexp(logpq, dim=0) - np.log(logp.shape[0]))
        return (torch.exp(logpq) * logpq).mean()

    @staticmethod
    def calc_kl_mean_includelogz(logq, logp):
        logqp = logq - logp
        return logqp.mean() + torch.logsumexp(-logqp, dim=0) - np.log(logp.shape[0])

    @staticmethod
    def calc_least_squares(logq, logp):
        logqp = logq - logp
        logz = torch.logsumexp(-logqp, dim=0) - np.log(logp.shape[0])
        return torch.mean((logqp + logz) ** 2)

    @staticmethod
    def calc_minus_logz(logq, logp):
        return -(torch.logsumexp(logp - logq, dim=0) - np.log(logp.shape[0]))

    @staticmethod
    def calc_ess(logq, logp):
        """Effective sample size, normalised to (0, 1]."""
        logqp = logq - logp
        log_ess = 2 * torch.logsumexp(-logqp, dim=0) - torch.logsumexp(-2 * logqp, dim=0)
        return torch.exp(log_ess) / len(logqp)

    def calc_minus_ess(self, logq, logp):
        return -self.calc_ess(logq, logp)

    @torch.no_grad()
    def _append_to_train_history(self, logq, logp):
        logqp = logq - logp
        hist = self.train_history
        hist['logz'].append(estimate_logz(logqp, method='jackknife'))
        hist['accept_rate'].append(self._model.mcmc.estimate_accept_rate(logqp))
        hist['ess'].append(self.calc_ess(logqp, 0))
        hist['rho'].append(self.calc_corrcoef(logq, logp))
        hist['logqp'].append((logqp.mean().item(), logqp.std().item()))

    def print_fit_status(self, epoch, loss=None):
        hist = self.train_history
        if loss is None:
            loss = hist['loss'][-1]
        logqp_mean, logqp_std = hist['logqp'][-1]
        logz_mean, logz_std = hist['logz'][-1]
        rate_mean, rate_std = hist['accept_rate'][-1]
        ess, rho = hist['ess'][-1], hist['rho'][-1]

        if epoch == 1:
            print(f"\n>>> Training progress ({ess.device}) <<<\n")
            print("Note: log(q/p) is estimated with normalized p; "
                  "mean & error are obtained from samples in a batch\n")

        epoch += self.checkpoint_dict['epochs_run']
        line = f"Epoch: {epoch} | loss: {loss:g} | ess: {ess:g} | rho: {rho:g}"
        line += " | log(z): {0} | log(q/p): {1} | accept_rate: {2}".format(
            fmt_val_err(logz_mean, logz_std, err_digits=2),
            fmt_val_err(logqp_mean + logz_mean, logqp_std, err_digits=2),   # p normalised by the estimate
            fmt_val_err(rate_mean, rate_std, err_digits=1))
        if self.checkpoint_dict['print_extra_func'] is not None:
            line += self.checkpoint_dict['print_extra_func'](epoch)
        print(line)


@torch.no_grad()
def backward_sanitychecker(model, n_samples=5, net_=None, return_details=False):
    """net_.backward(net_(x)) must give back x with a vanishing residual log-Jacobian
    (reference _normflowcore.py:432-451).  Prints sum|x - x_hat| and sum|log0_hat|."""
    if net_ is None:
        net_ = model.net_
    x = model.prior.sample(n_samples)
    y, logJ = net_(x)
    x_hat, log0_hat = net_.backward(y, log0=logJ)
    print("Sanity check is OK if following numbers are zero up to round off:")
    print(f"{torch.sum(torch.abs(x - x_hat)).item():g}", f"{torch.sum(torch.abs(log0_hat)).item():g}")
    if return_details:
        return (x, y, x_hat), (logJ, log0_hat)
