"""ctypes binding of libnormflow_b200.so (the C ABI declared in include/normflow_b200.h).

There is NO fallback: if the library is missing or a tensor is not a CUDA float32
tensor, the call raises.  PyTorch is used only for device memory and streams.
"""

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFK_LIB") or os.path.join(_PKG, "lib", "libnormflow_b200.so")   # NFK_LIB: tuning builds

c_f = ctypes.c_void_p      # every device pointer crosses as void*
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_fl = ctypes.c_float


class Lattice(ctypes.Structure):
    """nfk_lattice"""
    _fields_ = [("ndim", ctypes.c_int32), ("shape", ctypes.c_int32 * 4)]


class RqsParams(ctypes.Structure):
    """nfk_rqs_params"""
    _fields_ = [("n_knots", ctypes.c_int32), ("xlim0", c_fl), ("xlim1", c_fl),
                ("ylim0", c_fl), ("ylim1", c_fl),
                ("extrap_left", ctypes.c_int32), ("extrap_right", ctypes.c_int32)]


EXTRAP = {None: 0, 'none': 0, 'linear': 1, 'anti': 2, 'anti-periodic': 2, 'periodic': 3}
ACT = {None: 0, 'none': 0, 'tanh': 1, 'relu': 2, 'leaky_relu': 3, 'softplus': 4, 'abs': 5}
FROZEN_ZERO, FROZEN_COPY = 0, 1
EUNSUPPORTED = -2          # NFK_EUNSUPPORTED: the entry declines this configuration

# name -> argument types (return type is always int unless stated)
_SIGNATURES = {
    "nfk_mask_evenodd": [c_f, Lattice, c_i, c_i, c_f],
    "nfk_mask_alongaxis": [c_f, Lattice, c_i, c_i, c_f],
    "nfk_mask_select": [c_f, c_f, c_i, c_f, c_l, c_l, c_f],
    "nfk_prior_normal_sample": [c_f, c_f, c_l, c_l, c_f, c_f, c_u64, c_u64, c_f],
    "nfk_prior_normal_sample_dev": [c_f, c_f, c_l, c_l, c_f, c_f, c_f, c_f],
    "nfk_prior_normal_logprob": [c_f, c_f, c_l, c_l, c_f, c_f, c_f],
    "nfk_affine_fwd": [c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_affine_inv": [c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_affine_bwd": [c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_shift_apply": [c_f, c_f, c_f, c_i, c_i, c_fl, c_f, c_l, c_l, c_f],
    "nfk_rqs_fwd": [c_f, c_f, c_f, c_i, c_i, RqsParams, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_rqs_inv": [c_f, c_f, c_f, c_i, c_i, RqsParams, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_rqs_bwd": [c_f, c_f, c_f, c_i, c_i, RqsParams, c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_spline1d_fwd": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_spline1d_bwd": [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_logistic_fwd": [c_f, c_i, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_logistic_bwd": [c_f, c_i, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_phi4_action_fwd": [c_f, Lattice, c_fl, c_fl, c_fl, c_f, c_l, c_f],
    "nfk_phi4_action_bwd": [c_f, Lattice, c_fl, c_fl, c_fl, c_f, c_f, c_l, c_f],
    "nfk_conv_circ_fwd": [c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_f, c_i, c_f, Lattice, c_i, c_i, c_i, c_l, c_f],
    "nfk_conv_circ_fwd_cb": [c_f, c_i, c_f, c_i, c_f, c_i, c_f, c_i, c_f, Lattice, c_i, c_i, c_i, c_l, c_f],
    "nfk_conv_circ_bwd_weight": [c_f, c_f, c_i, c_f, c_f, c_f, Lattice, c_i, c_i, c_i, c_l, c_f],
    "nfk_conv_circ_bwd_weight_cb": [c_f, c_f, c_i, c_f, c_f, Lattice, c_i, c_i, c_i, c_l, c_f],
    "nfk_conv2d_wgrad_tc": [c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_l, c_f],
    "nfk_metropolis_scan": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_l, c_f],
    "nfk_metropolis_rates": [c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_gather_rows": [c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_fused2d_step": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, RqsParams, c_i, c_i, c_i,
                         c_f, c_f, c_f, c_i, c_i, c_l, c_f],
    "nfk_fused2d_step_train": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, RqsParams, c_i, c_i,
                               c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_l, c_f],
    "nfk_fusednd_step": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, RqsParams, Lattice, c_i, c_i, c_i,
                         c_f, c_f, c_f, c_l, c_f, c_l, c_f],
    "nfk_fusednd_step_train": [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, RqsParams, Lattice, c_i, c_i,
                               c_f, c_f, c_f, c_f, c_f, c_f, c_l, c_f, c_l, c_f],
    "nfk_convnd_dgrad": [c_f, c_i, c_f, c_f, c_f, c_i, c_i, Lattice, c_l, c_f, c_l, c_f],
    "nfk_convnd_wgrad": [c_f, c_f, c_f, c_f, c_i, c_i, Lattice, c_l, c_f, c_l, c_f],
    "nfk_convnd_layer_bwd": [c_f, c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_i, Lattice, c_l, c_f, c_l, c_f],
    "nfk_psd_weights_fwd": [c_f, c_l, c_i, c_i, c_f, c_f, c_f],
    "nfk_psd_weights_bwd": [c_f, c_f, c_f, c_f, c_l, c_i, c_i, c_f, c_f],
    "nfk_psd_scale": [c_f, c_f, c_f, c_fl, c_f, c_l, c_l, c_f],
    "nfk_psd_scale_bwd": [c_f, c_f, c_f, c_i, c_fl, c_f, c_f, c_f, c_f, c_l, c_l, c_f],
    "nfk_psd_chunks": [c_l, c_l],
    "nfk_knots_fwd": [c_f, c_f, c_f, c_i, c_fl, c_fl, c_fl, c_fl, c_f, c_f],
    "nfk_knots_bwd": [c_f, c_f, c_f, c_i, c_fl, c_fl, c_fl, c_fl, c_f, c_f, c_f, c_f, c_f],
    "nfk_sample_mean": [c_f, c_l, c_l, c_fl, c_f, c_f],
    "nfk_sample_shift": [c_f, c_f, c_f, c_l, c_l, c_f],
}

_lib = None


def declared_symbols():
    """Every entry point of include/normflow_b200.h (checked by the CPU test-suite)."""
    return sorted(list(_SIGNATURES) + ["nfk_strerror", "nfk_version", "nfk_launch_count", "nfk_fusednd_workspace",
                                            "nfk_convnd_dgrad_workspace", "nfk_convnd_wgrad_workspace",
                                            "nfk_convnd_layer_bwd_workspace"])


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"normflow_b200: {LIB_PATH} is missing -- build it with "
            "`python -m normflow__b200._build` (there is no CPU/eager fallback).")
    handle = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = c_i
    handle.nfk_fusednd_workspace.argtypes = [Lattice, c_i, c_i, c_i, c_l]
    handle.nfk_fusednd_workspace.restype = c_l
    handle.nfk_convnd_dgrad_workspace.argtypes = [Lattice, c_i, c_i, c_l]
    handle.nfk_convnd_dgrad_workspace.restype = c_l
    handle.nfk_convnd_wgrad_workspace.argtypes = [Lattice, c_i, c_i, c_l]
    handle.nfk_convnd_wgrad_workspace.restype = c_l
    handle.nfk_convnd_layer_bwd_workspace.argtypes = [Lattice, c_i, c_i, c_l]
    handle.nfk_convnd_layer_bwd_workspace.restype = c_l
    handle.nfk_strerror.argtypes = [c_i]
    handle.nfk_strerror.restype = ctypes.c_char_p
    handle.nfk_version.restype = c_i
    handle.nfk_launch_count.restype = ctypes.c_uint64
    _lib = handle
    return handle


def launch_count():
    return int(lib().nfk_launch_count())


def check(code, what):
    if code != 0:
        msg = lib().nfk_strerror(code).decode()
        raise RuntimeError(f"normflow_b200: {what} failed: {msg} ({code})")


def lattice(shape):
    shape = tuple(int(v) for v in shape)
    if not 1 <= len(shape) <= 4:
        raise ValueError(f"lattice dimension must be 1..4, got shape {shape}")
    return Lattice(len(shape), (ctypes.c_int32 * 4)(*(shape + (1,) * (4 - len(shape)))))


def dev(t, dtype=torch.float32, name="tensor"):
    """Validated device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"normflow_b200: {name} must live on a CUDA device "
                           "(the hot path has no CPU implementation)")
    if t.dtype != dtype:
        raise TypeError(f"normflow_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"normflow_b200: {name} must be contiguous")
    if t.get_device() != torch._C._cuda_getDevice():
        # the op wrappers (_ops._native) switch to the tensors' device; a pointer of another GPU on
        # the current device's stream would be an illegal address on the device, so refuse it here
        raise RuntimeError(f"normflow_b200: {name} lives on cuda:{t.get_device()} but the current device is "
                           f"cuda:{torch._C._cuda_getDevice()}: all tensors of one call must share a device")
    return t.data_ptr()


def stream():
    """Raw handle of torch's current CUDA stream on the current device (the C-level getters:
    `torch.cuda.current_stream()` costs ~15 us of Python per call)."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


# ---------------------------------------------------------------------------
# Optional per-kernel timing (bench.py): CUDA events recorded on the launching
# stream around selected launches.  Off by default (no overhead beyond a None test).
class KernelTimer:
    """Collects (start, stop) CUDA event pairs per kernel name; `summary()` syncs once."""

    def __init__(self):
        self.events = {}

    def record(self, name):
        return _Span(self, name)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, pairs in self.events.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            out[name] = dict(launches=len(ms), total_ms=float(sum(ms)), avg_ms=float(sum(ms) / max(len(ms), 1)))
        return out


class _Span:
    def __init__(self, timer, name):
        self.timer, self.name = timer, name

    def __enter__(self):
        self.start = torch.cuda.Event(enable_timing=True)
        self.stop = torch.cuda.Event(enable_timing=True)
        self.start.record()

    def __exit__(self, *exc):
        self.stop.record()
        self.timer.events.setdefault(self.name, []).append((self.start, self.stop))
        return False


class _NoSpan:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOSPAN = _NoSpan()
kernel_timer = None


def timed(name):
    return _NOSPAN if kernel_timer is None else kernel_timer.record(name)
