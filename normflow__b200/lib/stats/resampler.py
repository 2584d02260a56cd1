"""Bootstrap / jackknife / shuffling resampling (reference src/lib/stats/resampler.py)."""

import numpy as np
import torch


class Resampler:
    """Iterate over resampled copies of `samples` (axis 0).

    method: 'bootstrap' (draw with replacement, `n_resamples` times),
            'jackknife' (leave one bin out; one resample per bin),
            'shuffling' (random permutations, `n_resamples` times).
    """

    METHODS = ('bootstrap', 'jackknife', 'shuffling')

    def __init__(self, method='bootstrap'):
        if method not in self.METHODS:
            raise ValueError(f"unknown resampling method {method!r}")
        self.method = method

    def _index_sets(self, n_bins, n_resamples, batch_size, on_torch, device):
        if self.method == 'jackknife':
            everything = torch.arange(n_bins, device=device) if on_torch else np.arange(n_bins)
            for i in range(n_bins):
                yield everything[everything != i]
        elif self.method == 'bootstrap':
            size = n_bins if batch_size is None else batch_size
            for _ in range(n_resamples):
                yield (torch.randint(n_bins, size=(size,), device=device) if on_torch
                       else np.random.randint(n_bins, size=(size,)))
        else:
            for _ in range(n_resamples):
                yield torch.randperm(n_bins, device=device) if on_torch else np.random.permutation(n_bins)

    def __call__(self, samples, n_resamples=100, binsize=1, batch_size=None):
        on_torch = isinstance(samples, torch.Tensor)
        n_bins = samples.shape[0] // binsize
        binned = samples[:n_bins * binsize].reshape(n_bins, binsize, -1)
        device = samples.device if on_torch else None
        for ind in self._index_sets(n_bins, n_resamples, batch_size, on_torch, device):
            picked = binned[ind]
            yield picked.reshape(picked.shape[0] * binsize, *samples.shape[1:])

    def eval(self, samples, fn=lambda x: np.mean(x), **kwargs):
        """(mean, std) of `fn` over the resamples."""
        vals = [fn(q) for q in self(samples, **kwargs)]
        return np.mean(vals), np.std(vals)
