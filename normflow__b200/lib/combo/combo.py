"""log Z estimate and value(error) formatting (reference src/lib/combo/combo.py)."""

import numpy as np
import torch

from ..stats.resampler import Resampler


def estimate_logz(logqp, n_resamples=10, method='bootstrap'):
    """(mean, std) of log z from logqp = log q - log(p z), using
    z = E_q[exp(-logqp)]  (reference combo.py:11-23).

    The jackknife spread is evaluated in closed form -- leaving sample i out of a
    log-sum-exp L gives L + log(1 - exp(x_i - L)) -- so it costs one pass and one
    device-to-host copy instead of N resamples of N-1 elements with a sync each.
    """
    n = logqp.shape[0]
    if not isinstance(logqp, torch.Tensor):
        logqp = torch.as_tensor(np.asarray(logqp))
    x = -logqp.detach().double().reshape(-1)
    total = torch.logsumexp(x, dim=0)
    mean = total.item() - np.log(n)
    if method == 'jackknife':
        left_out = total + torch.log1p(-torch.exp(x - total).clamp(max=1 - 1e-15)) - np.log(n)
        std = left_out.std(unbiased=False).item()
    else:
        resampler = Resampler(method)
        vals = [torch.logsumexp(r.reshape(-1), dim=0).item() - np.log(n) for r in resampler(x, n_resamples)]
        std = float(np.std(vals))
    return mean, std


def fmt_val_err(value, error, err_digits=1):
    """1.2345, 0.0067 -> '1.2345(67)' style string (reference combo.py:26-34)."""
    try:
        if not error > 0:
            raise ValueError("no positive error to set the number of digits")
        digits = max(0, -int(np.floor(np.log10(error))) + err_digits - 1)
        return "{0:.{2}f}({1:.0f})".format(value, error * 10 ** digits, digits)
    except (ValueError, OverflowError, ZeroDivisionError):
        return "{0}+-{1}".format(value, error)
