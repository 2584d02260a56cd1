"""Device defaults and the multi-GPU handler (reference src/device/__init__.py:7-13).

Like the reference, importing the package makes CUDA the default device when one is
visible.  Unlike the reference the default dtype stays float32: the kernels compute in
fp32, which holds the required 1e-5 parity with the reference's float64 results
(hard part 1).
"""
import torch

from ._core import ModelDeviceHandler  # noqa: F401

torch_device = 'cuda' if torch.cuda.is_available() else 'cpu'
if torch_device == 'cuda':
    torch.set_default_device(torch_device)
