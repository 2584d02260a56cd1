"""Independence-Metropolis sampling on top of the flow (reference src/mcmc/mcmc.py).

The reference copies log q - log p to the host, walks the chain in a python loop,
copies indices back and gathers with three index_selects, with two `.item()` syncs per
batch (mcmc.py:55-87).  Here only the B uniforms (drawn and logged on the host with
numpy, exactly as the reference does, so decisions match it) travel to the device; the
sequential scan, the index construction and the row gather are kernels and the chain
state stays on the device -- no device-to-host sync per batch.
"""

import copy

import numpy as np
import torch

from .. import _ops
from ..lib.combo import estimate_logz, fmt_val_err
from ..lib.stats import Resampler


def seize(var):
    return var.detach().cpu().numpy()


class MCMCSampler:
    """Markov chain of flow proposals with Metropolis accept/reject
    (reference MCMCSampler, mcmc.py:15-128)."""

    def __init__(self, model):
        self._model = model
        self.history = MCMCHistory()
        self._reset_chain()

    def _reset_chain(self):
        # chain state carried across calls: last sample, its logq / logp, and the
        # device-side {ref, has_ref} pair read and written by the scan kernel
        self._ref = dict(sample=None, logq=None, logp=None, logqp=None)
        self._ref_state = None

    @torch.no_grad()
    def sample(self, batch_size=1, **kwargs):
        return self.sample__(batch_size=batch_size, **kwargs)[0]

    @torch.no_grad()
    def sample_(self, batch_size=1, **kwargs):
        return self.sample__(batch_size=batch_size, **kwargs)[:2]

    @torch.no_grad()
    def sample__(self, batch_size=1, bookkeeping=False):
        """(y, logq, logp) after accept/reject: rejected proposals repeat the last
        accepted configuration."""
        y, logq, logp = self._model.posterior.sample__(batch_size=batch_size)
        if bookkeeping:
            self.history.bookkeeping(raw_logq=logq, raw_logp=logp)
        y, logq, logp = self._accept_reject_step(y, logq, logp, bookkeeping=bookkeeping)
        if bookkeeping:
            self.history.bookkeeping(logq=logq, logp=logp)
        return y, logq, logp

    @torch.no_grad()
    def _accept_reject_step(self, y, logq, logp, bookkeeping=False):
        B = logq.shape[0]
        device = logq.device
        if self._ref_state is None or self._ref_state.device != device:
            self._ref_state = torch.zeros(2, dtype=torch.float64, device=device)
        # the reference draws np.random.rand(B) and takes its log on the host (mcmc.py:312)
        log_u = torch.from_numpy(np.log(np.random.rand(B))).to(device, non_blocking=True)
        accept, idx, n_acc = _ops.metropolis_scan(logq, logp, log_u, self._ref_state)

        ref = self._ref
        y = _ops.gather_rows(y, idx, ref['sample'])
        logq = _ops.gather_rows(logq, idx, ref['logq'])
        logp = _ops.gather_rows(logp, idx, ref['logp'])

        # next call's chain state (device tensors; no .item())
        ref['sample'], ref['logq'], ref['logp'] = y[-1].clone(), logq[-1:].clone(), logp[-1:].clone()
        ref['logqp'] = self._ref_state[:1]

        self.history.bookkeeping(accept_rate=n_acc.double() / B)
        if bookkeeping:
            accept_seq = seize(accept).astype(bool)
            # host view in the reference's convention: before the first acceptance the
            # index points at row 0, which holds the previous chain state (mcmc.py:67-68)
            self.history.bookkeeping(accept_seq=accept_seq, accept_ind=np.maximum(seize(idx), 0))
        return y, logq, logp

    @torch.no_grad()
    def serial_sample_generator(self, n_samples, batch_size=16):
        """Yield chain samples one at a time (each with a leading batch axis of 1)."""
        for i in range(n_samples):
            k = i % batch_size
            if k == 0:
                y, logq, logp = self.sample__(batch_size)
            yield y[k].unsqueeze(0), logq[k].unsqueeze(0), logp[k].unsqueeze(0)

    @torch.no_grad()
    def calc_accept_rate(self, n_samples=1024, batch_size=None, n_resamples=10, method='shuffling'):
        """Acceptance rate (mean, std) from fresh raw samples."""
        if batch_size is None or batch_size > n_samples:
            batch_size = n_samples
        n_batches = int(np.ceil(n_samples / batch_size))
        chunks = []
        for _ in range(n_batches):
            _, logq, logp = self._model.posterior.sample__(batch_size=batch_size)
            chunks.append(logq.double() - logp.double())
        return self.estimate_accept_rate(torch.cat(chunks))

    @staticmethod
    @torch.no_grad()
    def estimate_accept_rate(logqp, n_resamples=10, method='shuffling'):
        """Acceptance rate (mean, std) of chains built from shuffled copies of logqp (mcmc.py:117-124).

        A CUDA tensor stays on the device: the `n_resamples` permutations come from torch.randperm on
        the device and the uniforms from np.random.rand on the host -- the generators, and the order
        of their calls, that the reference uses for a tensor argument -- and the chains run
        concurrently, one warp each (`nfk_metropolis_rates`); one small device-to-host read returns
        the rates.  (The reference walks every chain in a host loop over CUDA elements, one
        synchronisation per element.)  Host arrays take the reference's own host loop."""
        if isinstance(logqp, torch.Tensor) and logqp.is_cuda and method == 'shuffling' and logqp.dim() == 1:
            n = logqp.shape[0]
            perms, log_u = [], np.empty((n_resamples, n))
            for r in range(n_resamples):              # per resample: randperm, then rand (resampler.py:62-64, mcmc.py:312)
                perms.append(torch.randperm(n, device=logqp.device))
                log_u[r] = np.log(np.random.rand(n))
            rates = _ops.metropolis_rates(logqp.double().contiguous(), torch.stack(perms),
                                          torch.from_numpy(log_u).to(logqp.device))
            rates = rates.cpu().numpy()
            return np.mean(rates), np.std(rates)
        if isinstance(logqp, torch.Tensor):
            logqp = seize(logqp).astype(np.float64)
        rate = lambda v: np.mean(Metropolis.calc_accept_status(v))
        return Resampler(method).eval(logqp, fn=rate, n_resamples=n_resamples)

    def log_prob(self, y, action_logz=0):
        return -self._model.action(y) - action_logz


class BlockedMCMCSampler(MCMCSampler):
    """Metropolis chain whose proposals redraw ONE block of the prior variables at a time
    (reference BlockedMCMCSampler, mcmc.py:132-219): a sweep visits the n_blocks blocks in order,
    each proposal is a single-configuration flow evaluation, accepted against the running
    log q - log p with log-uniforms drawn on the host as the reference draws them.

    The chain is sequential by construction (one configuration, one decision at a time); what
    this version saves is the flow evaluation the reference repeats after every sweep -- the
    state left by a sweep is the last accepted proposal, whose (y, logq, logp) are already known.
    """

    @torch.no_grad()
    def sample__(self, batch_size=1, n_blocks=1, bookkeeping=False):
        prior, net_ = self._model.prior, self._model.net_
        if self._ref['sample'] is not None:
            x = net_.backward(self._ref['sample'].unsqueeze(0))[0]
            logqp_ref = self._ref['logqp']
        else:
            print("Starting from scratch & setting logqp_ref to None")
            x = prior.sample(1)
            logqp_ref = None
        nvar = prior.nvar
        if isinstance(n_blocks, int):
            block_len = nvar // n_blocks
            assert block_len * n_blocks == nvar
        else:
            block_len, n_blocks = nvar, 1
        prior.setup_blockupdater(block_len)

        cfgs = torch.empty((batch_size, *prior.shape), dtype=torch.float32, device=x.device)
        logq = torch.empty((batch_size,), dtype=torch.float32, device=x.device)
        logp = torch.empty((batch_size,), dtype=torch.float32, device=x.device)
        accept_seq = np.empty((batch_size, n_blocks), dtype=bool)
        current = None                      # (y, logq, logp) of the configuration x stands for
        evaluate = self._make_evaluator(x)
        for ind in range(batch_size):
            accept_seq[ind], logqp_ref, current = self.sweep(x, n_blocks, logqp_ref, current=current,
                                                             return_state=True, evaluate=evaluate)
            cfgs[ind:ind + 1], logq[ind:ind + 1], logp[ind:ind + 1] = current

        self._ref['sample'] = cfgs[-1].clone()
        self._ref['logq'], self._ref['logp'] = float(logq[-1]), float(logp[-1])
        self._ref['logqp'] = float(logq[-1].double() - logp[-1].double())
        self.history.bookkeeping(accept_rate=np.mean(accept_seq))
        if bookkeeping:
            self.history.bookkeeping(logq=logq, logp=logp)
            self.history.bookkeeping(accept_seq=accept_seq.ravel())
        return cfgs, logq, logp

    cuda_graph = True       # replay the single-configuration evaluation as one CUDA graph (CUDA tensors only)

    def _evaluate(self, x):
        y, logJ = self._model.net_(x)
        return y, self._model.prior.log_prob(x) - logJ, -self._model.action(y)

    def _make_evaluator(self, x):
        """The evaluation of the ONE configuration buffer `x`, as a callable.  A sweep is a chain of
        single-configuration flow evaluations, each a dozen or more launches of a few microseconds: captured
        once per `sample__` call and replayed, it is one graph launch per proposal.  The callable returns
        buffers that the next call overwrites."""
        if not (self.cuda_graph and x.is_cuda):
            return lambda: self._evaluate(x)
        try:
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(2):                   # first-use initialisation happens outside the capture
                    self._evaluate(x)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._evaluate(x)
        except Exception as err:                     # e.g. a user conditioner that cannot be captured
            print(f"BlockedMCMCSampler: CUDA graph capture failed ({type(err).__name__}); evaluating eagerly")
            torch.cuda.synchronize(x.device)
            return lambda: self._evaluate(x)

        def replay():
            graph.replay()
            return out
        return replay

    @torch.no_grad()
    def sweep(self, x, n_blocks=1, logqp_ref=None, current=None, return_state=False, evaluate=None):
        """One in-place sweep over the blocks of x (shape (1, *lattice)); returns the accept flags and
        the updated reference log q - log p (mcmc.py:196-219)."""
        prior = self._model.prior
        if evaluate is None:
            evaluate = lambda: self._evaluate(x)
        accept_seq = np.empty(n_blocks, dtype=bool)
        lrand_arr = np.log(np.random.rand(n_blocks))
        for ind in range(n_blocks):
            prior.blockupdater(x, ind)
            proposal = evaluate()
            logqp = float((proposal[1].double() - proposal[2].double())[0])      # the step's one host sync
            if ind == 0 and logqp_ref is None:
                accept_seq[ind] = True
            else:
                accept_seq[ind] = lrand_arr[ind] < logqp_ref - logqp
            if accept_seq[ind]:
                logqp_ref, current = logqp, tuple(t.clone() for t in proposal)     # the evaluator reuses its buffers
            else:
                prior.blockupdater.restore(x, ind)
        if not return_state:
            return accept_seq, logqp_ref
        if current is None:                 # every proposal rejected and no earlier evaluation at hand
            current = tuple(t.clone() for t in evaluate())
        return accept_seq, logqp_ref, current


class MCMCHistory:
    """Bookkeeping lists of a simulation (reference MCMCHistory, mcmc.py:223-294)."""

    def __init__(self):
        self.reset_history()

    def reset_history(self):
        self.logq, self.logp = [], []
        self.raw_logq, self.raw_logp = [], []
        self.accept_seq, self.accept_ind = [], []
        self._accept_rate = []

    @property
    def accept_rate(self):
        """Acceptance rate of every batch as python floats.  Entries are kept as device
        scalars until somebody looks, so sampling itself never waits for the GPU."""
        self._accept_rate = [float(a) for a in self._accept_rate]
        return self._accept_rate

    def bookkeeping(self, logq=None, logp=None, raw_logq=None, raw_logp=None, accept_seq=None,
                    accept_rate=None, accept_ind=None):
        if raw_logq is not None:
            self.raw_logq.append(copy.copy(seize(raw_logq)))
        if raw_logp is not None:
            self.raw_logp.append(copy.copy(seize(raw_logp)))
        if logq is not None:
            self.logq.append(seize(logq))
        if logp is not None:
            self.logp.append(seize(logp))
        if accept_rate is not None:
            self._accept_rate.append(accept_rate)
        if accept_seq is not None:
            self.accept_seq.append(accept_seq)
        if accept_ind is not None:
            self.accept_ind.append(accept_ind)

    def report_summary(self, since=0, asstr=False):
        fmt = (lambda m, s: fmt_val_err(m, s, err_digits=2)) if asstr else (lambda m, s: (m, s))
        logqp = torch.tensor(self.logq[-1] - self.logp[-1])
        rate = torch.tensor(self.accept_rate)
        stats = lambda t: (t.mean().item(), t.std().item())
        return {'logqp': fmt(*stats(logqp)), 'logz': fmt(*estimate_logz(logqp)),
                'accept_rate': fmt(*stats(rate))}

    @property
    def logqp(self):
        return [q - p for q, p in zip(self.logq, self.logp)]

    @property
    def raw_logqp(self):
        return [q - p for q, p in zip(self.raw_logq, self.raw_logp)]


class Metropolis:
    """Host-side Metropolis-Hastings helpers with the reference's signatures
    (mcmc.py:298-352); used for diagnostics on small host arrays.  The sampler's own
    accept/reject runs in `nfk_metropolis_scan`."""

    @staticmethod
    @torch.no_grad()
    def calc_accept_status(logqp, logqp_ref=None):
        """accept[i] = log u_i < ref - logqp_i, ref <- logqp_i on acceptance; consumes
        exactly one np.random.rand(len(logqp))."""
        logqp = np.asarray(logqp)
        if logqp_ref is None:
            logqp_ref = logqp[0]
        status = np.empty(len(logqp), dtype=bool)
        log_u = np.log(np.random.rand(logqp.shape[0]))
        for i in range(len(logqp)):
            status[i] = log_u[i] < (logqp_ref - logqp[i])
            if status[i]:
                logqp_ref = logqp[i]
        return status

    @staticmethod
    def calc_accept_indices(accept_seq):
        """index of the most recent accepted proposal (0 before the first one)."""
        accept_seq = np.asarray(accept_seq, dtype=bool)
        own = np.where(accept_seq, np.arange(len(accept_seq)), 0)
        return np.maximum.accumulate(own)

    @staticmethod
    def calc_accept_count(accept_seq):
        hits = np.where(accept_seq)[0]
        return hits[1:] - hits[:-1]

    @staticmethod
    def calc_tau_rejections_prob(accept_seq, max_tau=100):
        """Probability of tau+1 rejections in a row, tau = 0..max_tau-1."""
        rejected = ~np.asarray(accept_seq, dtype=bool)
        run = rejected
        p_tau = np.zeros(max_tau)
        p_tau[0] = np.mean(run)
        for tau in range(1, max_tau):
            run = run[:-1] & rejected[tau:]
            p_tau[tau] = np.mean(run) if len(run) else 0.0
        return p_tau
