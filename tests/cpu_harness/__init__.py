"""Builds and loads the host-side harness around the kernels' per-site functors.
Test infrastructure only (see harness.cpp)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "harness.cpp")
OUT = os.path.join(HERE, "_build", "harness.so")
DEPS = [SRC] + [os.path.join(ROOT, "normflow__b200", "csrc", f) for f in ("nfk_math.cuh", "nfk_ops.cuh", "nfk_knots.cuh", "nfk_psd.cuh")]


class Lattice(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int32), ("shape", ctypes.c_int32 * 4)]


class RqsParams(ctypes.Structure):
    _fields_ = [("n_knots", ctypes.c_int32), ("xlim0", ctypes.c_float), ("xlim1", ctypes.c_float),
                ("ylim0", ctypes.c_float), ("ylim1", ctypes.c_float),
                ("extrap_left", ctypes.c_int32), ("extrap_right", ctypes.c_int32)]


def lattice(shape):
    shape = tuple(int(v) for v in shape)
    return Lattice(len(shape), (ctypes.c_int32 * 4)(*(shape + (1,) * (4 - len(shape)))))


def load():
    if (not os.path.exists(OUT)) or max(os.path.getmtime(d) for d in DEPS) > os.path.getmtime(OUT):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    return ctypes.CDLL(OUT)
